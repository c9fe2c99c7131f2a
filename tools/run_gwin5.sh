#!/bin/bash
mkdir -p gpurun_out/gwin
timeout 300 python -m pytest tests/test_csr_gpu.py -x -q -m gpu -k "global or rmat or families or option" > gpurun_out/gwin/tests5.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/gwin/tests5.log
for s in 18 20; do
  timeout 300 python bench.py --workload rmat --scale $s --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/gwin/rmat${s}_v5.json 2> gpurun_out/gwin/rmat${s}_v5.err
done
IAS_OPT_GWIN_MAX_SW=0 timeout 300 python bench.py --workload rmat --scale 22 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/gwin/rmat22_v5.json 2> gpurun_out/gwin/rmat22_v5.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/gwin/*_v5.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); c=d['config']
            print(f.split('/')[-1], 'ms', round(d['ms_per_step'],2), 'GF', round(d['value'],1), 'sym', c['ms_bin_sym'], 'num', c['ms_bin_num'])
PY
tail -n 2 gpurun_out/gwin/*_v5.err
