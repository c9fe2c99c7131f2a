#!/bin/bash
# round 2: the default bench line on N GPUs of one box, launched exactly as the driver does
N=${1:-2}
O=gpurun_out/r02_n$N
mkdir -p $O
nvidia-smi --query-gpu=index,name,memory.used --format=csv > $O/gpus.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "bench N=$N rc=$?"; tail -5 $O/bench_n$N.err | cut -c1-300
python -c "
import json; d=json.load(open('$O/bench_n$N.json'))
print('main', d['n_gpus'], d['config']['format'], round(d['ms_per_step'],3), round(d['value'],1), d['scaling'], 'e2e', d['e2e'])
for k,v in d.get('also',{}).items(): print(k, json.dumps(v)[:1500])
"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/bench_ref_n$N.json 2> $O/bench_ref_n$N.err
echo "ref N=$N rc=$?"; cut -c1-600 $O/bench_ref_n$N.json
