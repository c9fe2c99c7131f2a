#!/bin/bash
mkdir -p gpurun_out/gwin
timeout 300 python -m pytest tests/test_csr_gpu.py -x -q -m gpu -k "families or option or rmat" > gpurun_out/gwin/tests2.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/gwin/tests2.log
timeout 300 python bench.py --workload rmat --scale 18 --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/gwin/rmat18_smem.json 2> gpurun_out/gwin/rmat18_smem.err
IAS_LIB=$PWD/ia_spgemm_b200/libiaspgemm_prof.so timeout 300 python bench.py --workload rmat --scale 20 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/gwin/rmat20_prof.json 2> gpurun_out/gwin/rmat20_prof.err
IAS_LIB=$PWD/ia_spgemm_b200/libiaspgemm_prof.so timeout 300 python bench.py --workload rmat --scale 18 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/gwin/rmat18_prof.json 2> gpurun_out/gwin/rmat18_prof.err
grep -h "gwin" gpurun_out/gwin/rmat20_prof.err | tail -4
grep -h "gwin" gpurun_out/gwin/rmat18_prof.err | tail -2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/gwin/rmat18_smem.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); c=d['config']
            print(f.split('/')[-1], 'ms', round(d['ms_per_step'],2), 'GF', round(d['value'],1), 'sym', c['ms_bin_sym'], 'num', c['ms_bin_num'])
PY
tail -n 3 gpurun_out/gwin/rmat18_smem.err
