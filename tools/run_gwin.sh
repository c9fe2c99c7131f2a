#!/bin/bash
# windowed global-row kernels: parity tests, then A/B against the L2 bitmap kernels on R-MAT
mkdir -p gpurun_out/gwin
timeout 600 python -m pytest tests/test_csr_gpu.py -x -q -m gpu > gpurun_out/gwin/tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/gwin/tests.log
for s in 18 20; do
  IAS_OPT_GLOBAL_ROWS_SMEM=0 timeout 300 python bench.py --workload rmat --scale $s --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/gwin/rmat${s}_l2.json 2> gpurun_out/gwin/rmat${s}_l2.err
  timeout 300 python bench.py --workload rmat --scale $s --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/gwin/rmat${s}_smem.json 2> gpurun_out/gwin/rmat${s}_smem.err
done
timeout 400 python bench.py --workload rmat --scale 22 --no-cpu --no-e2e --steps 2 --warmup 2 > gpurun_out/gwin/rmat22_smem.json 2> gpurun_out/gwin/rmat22_smem.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/gwin/*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); c=d['config']
            print(f.split('/')[-1], 'ms', round(d['ms_per_step'],2), 'GF', round(d['value'],1), 'sym', c['ms_bin_sym'], 'num', c['ms_bin_num'])
PY
tail -3 gpurun_out/gwin/*.err | tail -30
