#!/bin/bash
mkdir -p gpurun_out/gwin
timeout 300 python -m pytest tests/test_csr_gpu.py -x -q -m gpu -k "global or rmat or families or option" > gpurun_out/gwin/tests9.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/gwin/tests9.log
for gc in 0 2048 1048576; do
  IAS_OPT_G_CACHE=$gc timeout 300 python bench.py --workload rmat --scale 20 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/gwin/rmat20_c${gc}_v9.json 2> gpurun_out/gwin/rmat20_c${gc}_v9.err
done
timeout 300 python bench.py --workload rmat --scale 22 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/gwin/rmat22_v9.json 2> gpurun_out/gwin/rmat22_v9.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/gwin/*_v9.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); c=d['config']
            print(f.split('/')[-1], 'ms', round(d['ms_per_step'],2), 'GF', round(d['value'],1), 'sym', c['ms_bin_sym'][5], 'num', c['ms_bin_num'][5])
PY
tail -n 2 gpurun_out/gwin/*_v9.err
