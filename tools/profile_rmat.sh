#!/bin/bash
# ncu evidence for the R-MAT path. usage: profile_rmat.sh SCALE KERNEL_REGEX SKIP COUNT
mkdir -p gpurun_out
CMD="python bench.py --workload rmat --scale ${1:-16} --steps 1 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_r.json 2> gpurun_out/plain_r.err &&
ncu --set full --clock-control none --import-source on -k regex:"${2:-k_num_global|k_sym_global|k_num_hash_cta|k_sym_hash|k_esc_warp}" -s ${3:-27} -c ${4:-9} -o gpurun_out/prof_r $CMD > gpurun_out/ncu_full_r.log 2>&1
echo done
