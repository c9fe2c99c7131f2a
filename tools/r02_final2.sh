#!/bin/bash
# round 2, closing run on one GPU: the driver's sequence, then two ncu captures of the final kernels
O=gpurun_out/r02_final2
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q > $O/tests_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/tests_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-200 $O/bench_reference.json
timeout 1500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -2 $O/bench_default.err | cut -c1-200
python -c "
import json; d=json.load(open('$O/bench_default.json'))
print('main', d['config']['format'], round(d['ms_per_step'],3), round(d['value'],1), 'frac', round(d['roofline']['frac'],3), 'e2e', d['e2e'])
for k,v in d.get('also',{}).items(): print(k, json.dumps({a:b for a,b in v.items() if a not in ('roofline','detail')})[:700])
print('cpu', d['cpu_baseline'], 'clocks', d['clocks'])
"
B="python bench.py --no-also --no-cpu --no-e2e --no-cusparse --steps 2 --warmup 3"
$B --workload rmat --scale 20 > $O/plain_rmat20.json 2> $O/plain_rmat20.err &&
ncu --set full --clock-control none --import-source on -k regex:k_num_global2 -s 0 -c 1 -o $O/global2 $B --workload rmat --scale 20 > $O/ncu_global2.log 2>&1; echo "global2 rc=$?"
$B --workload uniform --format ell > $O/plain_uniform.json 2> $O/plain_uniform.err &&
ncu --set full --clock-control none --import-source on -k regex:k_ell_mul_ell -s 3 -c 1 -o $O/ell $B --workload uniform --format ell > $O/ncu_ell.log 2>&1; echo "ell rc=$?"
