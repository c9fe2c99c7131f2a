#!/bin/bash
# ncu evidence for the uniform 8M x 16 workload (warp bin). Run under gpurun, 1 GPU.
mkdir -p gpurun_out
CMD="python bench.py --workload uniform --steps 1 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_u.json 2> gpurun_out/plain_u.err &&
ncu --set full --clock-control none --import-source on -k regex:"k_esc_warp|k_sym_hash" -s 6 -c 2 -o gpurun_out/prof_u $CMD > gpurun_out/ncu_full_u.log 2>&1
echo done
