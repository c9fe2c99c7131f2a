#!/bin/bash
# Poisson weak-scaling line at N GPUs (what the driver's SCALE run launches)
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_p_n$N.json 2> gpurun_out/bench_p_n$N.err
echo "poisson N=$N rc=$?"
