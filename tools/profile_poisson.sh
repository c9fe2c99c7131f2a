#!/bin/bash
# ncu evidence for the default bench workload (Poisson 4096^2, CSR path). Run under gpurun, 1 GPU.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_p.json 2> gpurun_out/plain_p.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_p.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_num_tiny|k_sym_tiny|k_row_ub_thread|k_classify_num" -s 9 -c 3 -o gpurun_out/prof_p $CMD > gpurun_out/ncu_full.log 2>&1
echo done
