#!/bin/bash
# round 2, 2-GPU call: row-list test, then the default line on N GPUs as the driver launches it
N=${1:-2}
O=gpurun_out/r02_n${N}b
mkdir -p $O
timeout 600 python -m pytest tests/test_csr_gpu.py -m gpu -x -q --tb=short -k "row_list or stream" > $O/test_rowlist.log 2>&1; echo "rowlist tests rc=$? $(tail -1 $O/test_rowlist.log)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "bench N=$N rc=$?"; tail -3 $O/bench_n$N.err | cut -c1-300
python -c "
import json; d=json.load(open('$O/bench_n$N.json'))
print('main', d['n_gpus'], d['config']['format'], round(d['ms_per_step'],3), round(d['value'],1), d['scaling'], 'e2e', d['e2e'])
for k,v in d.get('also',{}).items(): print(k, json.dumps({a:b for a,b in v.items() if a not in ('roofline','detail')})[:1200])
"
