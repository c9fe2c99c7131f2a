#!/bin/bash
# round 2: the device block cache -- stalls gone?  (host trace of bench.py's own order of legs, then the default bench)
O=gpurun_out/r02_cache; mkdir -p $O
IAS_HOST_TRACE=1 timeout 120 python tools/stall_probe2.py > $O/probe2_trace.log 2>&1; echo "probe rc=$?"; grep -c "host trace" $O/probe2_trace.log; grep -v "host trace" $O/probe2_trace.log | tail -8
IAS_OPT_BLOCK_CACHE=0 IAS_HOST_TRACE=1 timeout 120 python tools/stall_probe2.py > $O/probe2_trace_nocache.log 2>&1; echo "probe(nocache) rc=$?"; grep -c "host trace" $O/probe2_trace_nocache.log; grep -v "host trace" $O/probe2_trace_nocache.log | tail -8
timeout 1200 python -m pytest tests -m gpu -q -x --durations=15 --deselect tests/test_fullsize_gpu.py > $O/tests_gpu_part.log 2>&1; echo "pytest rc=$? $(tail -1 $O/tests_gpu_part.log)"
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -2 $O/bench_default.err | cut -c1-200
python -c "
import json; d=json.load(open('$O/bench_default.json'))
print('main', d['config']['format'], round(d['ms_per_step'],3), round(d['value'],1), 'frac', round(d['roofline']['frac'],3), 'e2e', d['e2e'])
for k,v in d.get('also',{}).items(): print(k, json.dumps({a:b for a,b in v.items() if a not in ('roofline','detail')})[:500])
print('cpu', d['cpu_baseline'], 'clocks', d['clocks'])
"
