#!/bin/bash
mkdir -p gpurun_out/gwin
for s in 20 22; do
IAS_LIB=$PWD/ia_spgemm_b200/libiaspgemm_prof.so timeout 300 python bench.py --workload rmat --scale $s --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/gwin/rmat${s}_prof4.json 2> gpurun_out/gwin/rmat${s}_prof4.err
done
grep -h "gwin num" gpurun_out/gwin/rmat20_prof4.err | tail -3
echo ---
grep -h "gwin num" gpurun_out/gwin/rmat22_prof4.err | tail -19
