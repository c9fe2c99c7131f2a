#!/bin/bash
mkdir -p gpurun_out/gwin
timeout 300 python -m pytest tests/test_csr_gpu.py -x -q -m gpu -k "global or rmat or families or option or row_blocks or streaming" > gpurun_out/gwin/tests8.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/gwin/tests8.log
timeout 300 python bench.py --workload rmat --scale 18 --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/gwin/rmat18_v8.json 2> gpurun_out/gwin/rmat18_v8.err
IAS_OPT_GWIN_SWORDS=4096 timeout 300 python bench.py --workload rmat --scale 18 --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/gwin/rmat18h_v8.json 2> gpurun_out/gwin/rmat18h_v8.err
for s in 20 22; do
  timeout 300 python bench.py --workload rmat --scale $s --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/gwin/rmat${s}_v8.json 2> gpurun_out/gwin/rmat${s}_v8.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/gwin/*_v8.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); c=d['config']
            print(f.split('/')[-1], 'ms', round(d['ms_per_step'],2), 'GF', round(d['value'],1), 'sym', c['ms_bin_sym'], 'num', c['ms_bin_num'])
PY
tail -n 2 gpurun_out/gwin/*_v8.err
