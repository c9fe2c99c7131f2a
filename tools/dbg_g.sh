#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 4; do
  IAS_G_DBG=$d python bench.py --workload rmat --scale 20 --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/dbg_$d.json 2> gpurun_out/dbg_$d.err
done
