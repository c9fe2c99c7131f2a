#!/bin/bash
mkdir -p gpurun_out/gwin
IAS_LIB=$PWD/ia_spgemm_b200/libiaspgemm_prof.so timeout 300 python bench.py --workload rmat --scale 20 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/gwin/rmat20_prof6.json 2> gpurun_out/gwin/rmat20_prof6.err
grep -h "gwin num" gpurun_out/gwin/rmat20_prof6.err | tail -3
