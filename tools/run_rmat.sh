#!/bin/bash
mkdir -p gpurun_out
for s in "$@"; do
  python bench.py --workload rmat --scale $s --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/bench_r$s.json 2> gpurun_out/bench_r$s.err
  tail -2 gpurun_out/bench_r$s.err
done
