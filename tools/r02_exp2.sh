#!/bin/bash
# round 2, GPU call 2: whole GPU test-suite, smoke, the new default bench line (N=1) and the reference arm
O=gpurun_out/r02_exp2
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q > $O/tests_gpu.log 2>&1; echo "pytest rc=$?" >> $O/tests_gpu.log
tail -15 $O/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
timeout 1500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -5 $O/bench_default.err
python -c "
import json; d=json.load(open('$O/bench_default.json'))
print('main', d['config']['format'], round(d['ms_per_step'],3), round(d['value'],1), 'frac', round(d['roofline']['frac'],3), 'e2e', d['e2e'])
for k,v in d.get('also',{}).items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms_per_step','value','error','selected_format')}, v.get('cusparse'), v.get('ell_path',{}).get('ms_per_step'), v.get('e2e'))
print('cpu', d['cpu_baseline'])
"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cat $O/bench_reference.json | cut -c1-400
