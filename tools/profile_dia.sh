#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --format dia --steps 1 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_d.json 2> gpurun_out/plain_d.err &&
ncu --set full --clock-control none --import-source on -k regex:"k_dia_mul_dia" -s 3 -c 1 -o gpurun_out/prof_d $CMD > gpurun_out/ncu_full_d.log 2>&1
echo done
