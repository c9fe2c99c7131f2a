#!/bin/bash
# Round-1 closing evidence, one GPU: full parity suite, smoke, the bench lines quoted in profiles/r01_summary.md,
# and the ncu captures of the global-row kernels.  Outputs under gpurun_out/final2/.
O=gpurun_out/final2
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/tests_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/poisson_reference.json 2> $O/poisson_reference.err
timeout 600 python bench.py --steps 10 --warmup 3 > $O/poisson_csr.json 2> $O/poisson_csr.err
for s in 16 18 20 22; do
  timeout 400 python bench.py --workload rmat --scale $s --no-cpu --no-e2e --steps 2 --warmup 3 > $O/rmat$s.json 2> $O/rmat$s.err
done
# ncu: windowed kernels at scale 18 (launch list of one step + full capture), L2 kernel with the shared-memory mark at scale 20
CMD18="python bench.py --workload rmat --scale 18 --steps 1 --warmup 3 --no-cpu --no-e2e"
CMD20="python bench.py --workload rmat --scale 20 --steps 1 --warmup 3 --no-cpu --no-e2e"
$CMD18 > $O/plain_r18.json 2> $O/plain_r18.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file $O/launches_r18.csv $CMD18 > $O/ncu_list_r18.log 2>&1
$CMD18 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_num_gwin|k_sym_gwin" -s 6 -c 2 -o $O/prof_r18 $CMD18 > $O/ncu_full_r18.log 2>&1
$CMD20 > $O/plain_r20.json 2> $O/plain_r20.err &&
ncu --set full --clock-control none --import-source on -k regex:"k_num_global" -s 9 -c 1 -o $O/prof_r20 $CMD20 > $O/ncu_full_r20.log 2>&1
ls -la $O | tail -30
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/final2/*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); c=d.get('config',{})
            print(f.split('/')[-1], 'ms', round(d['ms_per_step'],2), d['unit'], round(d['value'],1), (d.get('roofline') or {}).get('kernel'), round((d.get('roofline') or {}).get('frac',0),3), (d.get('e2e') or {}).get('value'))
PY
