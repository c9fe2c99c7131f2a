#!/bin/bash
# round 2: the driver's sequence on one GPU -- whole GPU suite, smoke, reference arm, default bench line
O=gpurun_out/r02_final
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q > $O/tests_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/tests_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-300 $O/bench_reference.json
timeout 1500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -2 $O/bench_default.err | cut -c1-200
python -c "
import json; d=json.load(open('$O/bench_default.json'))
print('main', d['config']['format'], round(d['ms_per_step'],3), round(d['value'],1), 'frac', round(d['roofline']['frac'],3), 'e2e', d['e2e'])
for k,v in d.get('also',{}).items(): print(k, json.dumps({a:b for a,b in v.items() if a not in ('roofline','detail')})[:900])
print('cpu', d['cpu_baseline'], 'clocks', d['clocks'])
"
