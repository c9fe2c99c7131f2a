#!/bin/bash
# round 2, closing run on one GPU: the driver's sequence (GPU tests, smoke, reference arm, default bench), then launch lists
O=gpurun_out/r02_close; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/tests_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/tests_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-160 $O/bench_reference.json
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$O/bench_default.json'))
print('main', d['config']['format'], round(d['ms_per_step'],3), round(d['value'],1), 'frac', round(d['roofline']['frac'],3), 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['per_call_ms'])
for k,v in d.get('also',{}).items(): print(k, json.dumps({a:b for a,b in v.items() if a not in ('roofline','detail')})[:400])
print('cpu', d['cpu_baseline'], 'clocks', d['clocks'])
"
B="python bench.py --no-also --no-cpu --no-e2e --no-cusparse --steps 2 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_poisson_auto.csv $B > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_rmat20.csv $B --workload rmat --scale 20 > $O/ncu_launches_rmat20.log 2>&1; echo "launch list rmat20 rc=$?"
