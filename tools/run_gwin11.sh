#!/bin/bash
mkdir -p gpurun_out/gwin
timeout 120 python -m pytest tests/test_csr_gpu.py -x -q -m gpu > gpurun_out/gwin/tests11.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/gwin/tests11.log
IAS_OPT_G_L2_PERSIST=0 timeout 100 python bench.py --workload rmat --scale 20 --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/gwin/rmat20_p0_v11.json 2> gpurun_out/gwin/rmat20_p0_v11.err
timeout 100 python bench.py --workload rmat --scale 20 --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/gwin/rmat20_p1_v11.json 2> gpurun_out/gwin/rmat20_p1_v11.err
timeout 100 python bench.py --workload rmat --scale 22 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/gwin/rmat22_p1_v11.json 2> gpurun_out/gwin/rmat22_p1_v11.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/gwin/*_v11.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); c=d['config']
            print(f.split('/')[-1], 'ms', round(d['ms_per_step'],2), 'GF', round(d['value'],1), 'sym', c['ms_bin_sym'][5], 'num', c['ms_bin_num'][3:])
PY
tail -n 2 gpurun_out/gwin/*_v11.err
