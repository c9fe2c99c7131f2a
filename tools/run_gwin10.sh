#!/bin/bash
mkdir -p gpurun_out/gwin
timeout 300 python -m pytest tests/test_csr_gpu.py tests/test_formats_gpu.py -x -q -m gpu -k "global or rmat or families or option or every_bin or row_blocks or streaming" > gpurun_out/gwin/tests10.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/gwin/tests10.log
for s in 16 18; do
  IAS_OPT_GWIN_TAKES_B2=0 timeout 300 python bench.py --workload rmat --scale $s --no-cpu --no-e2e --steps 3 --warmup 3 > gpurun_out/gwin/rmat${s}_b2off_v10.json 2> gpurun_out/gwin/rmat${s}_b2off_v10.err
  timeout 300 python bench.py --workload rmat --scale $s --no-cpu --no-e2e --steps 3 --warmup 3 > gpurun_out/gwin/rmat${s}_v10.json 2> gpurun_out/gwin/rmat${s}_v10.err
done
IAS_OPT_G_COOP=0 timeout 300 python bench.py --workload rmat --scale 20 --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/gwin/rmat20_coopoff_v10.json 2> gpurun_out/gwin/rmat20_coopoff_v10.err
timeout 300 python bench.py --workload rmat --scale 20 --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/gwin/rmat20_v10.json 2> gpurun_out/gwin/rmat20_v10.err
timeout 300 python bench.py --workload rmat --scale 22 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/gwin/rmat22_v10.json 2> gpurun_out/gwin/rmat22_v10.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/gwin/*_v10.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); c=d['config']
            print(f.split('/')[-1], 'ms', round(d['ms_per_step'],2), 'GF', round(d['value'],1), 'sym', c['ms_bin_sym'][3:], 'num', c['ms_bin_num'][3:], c['num_bin_rows'][3:])
PY
tail -n 2 gpurun_out/gwin/*_v10.err
