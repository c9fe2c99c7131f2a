mkdir -p gpurun_out; python bench.py --steps 5 --warmup 3 > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; echo "rc=$?" >> gpurun_out/bench_p.err
python bench.py --format dia --no-cpu --steps 5 > gpurun_out/bench_dia.json 2> gpurun_out/bench_dia.err
python bench.py --workload uniform --n 8000000 --no-cpu --no-e2e --steps 3 > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err
python bench.py --workload rmat --scale 16 --no-cpu --no-e2e --steps 3 > gpurun_out/bench_r16.json 2> gpurun_out/bench_r16.err
python bench.py --workload rmat --scale 18 --no-cpu --no-e2e --steps 3 > gpurun_out/bench_r18.json 2> gpurun_out/bench_r18.err
