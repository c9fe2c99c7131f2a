#!/bin/bash
# every single-GPU number quoted in profiles/r01_summary.md, in one go
mkdir -p gpurun_out/final
python bench.py --steps 10 --warmup 3 > gpurun_out/final/poisson_csr.json 2> gpurun_out/final/poisson_csr.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/poisson_reference.json 2> gpurun_out/final/poisson_reference.err
python bench.py --format dia --no-cpu --steps 10 > gpurun_out/final/poisson_dia.json 2> gpurun_out/final/poisson_dia.err
python bench.py --workload uniform --steps 3 --no-e2e > gpurun_out/final/uniform_csr.json 2> gpurun_out/final/uniform_csr.err
python bench.py --workload uniform --format ell --steps 3 --no-cpu --no-e2e > gpurun_out/final/uniform_ell.json 2> gpurun_out/final/uniform_ell.err
for s in 16 18 20 22; do
  python bench.py --workload rmat --scale $s --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/final/rmat$s.json 2> gpurun_out/final/rmat$s.err
done
python bench.py --workload rmat --scale 18 --steps 2 --no-e2e > gpurun_out/final/rmat18_cpu.json 2> gpurun_out/final/rmat18_cpu.err
echo done
