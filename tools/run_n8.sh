#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_p_n$N.json 2> gpurun_out/bench_p_n$N.err
echo "poisson rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --workload rmat --scale ${2:-25} --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r${2:-25}_n$N.json 2> gpurun_out/bench_r${2:-25}_n$N.err
echo "rmat rc=$?"
tail -3 gpurun_out/bench_r${2:-25}_n$N.err
