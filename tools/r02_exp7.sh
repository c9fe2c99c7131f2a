#!/bin/bash
O=gpurun_out/r02_exp7
mkdir -p $O
timeout 900 python -m pytest tests/test_csr_gpu.py -m gpu -x -q --tb=short > $O/test_csr_gpu.log 2>&1; echo "test_csr_gpu rc=$? $(tail -1 $O/test_csr_gpu.log)"
timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -x -q --tb=short -k "rmat" > $O/test_fullsize_rmat.log 2>&1; echo "test_fullsize rmat rc=$? $(tail -1 $O/test_fullsize_rmat.log)"
run() {  # name, env..., -- args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu --no-also --no-cusparse --no-e2e --steps 2 --warmup 3 "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$O/$name.json')); print(d['config']['format'], round(d['ms_per_step'],3),'ms', round(d['value'],1),'GF', 'step_frac', round(d['roofline']['step_frac'],3), d['config'].get('ms_bin_sym'), d['config'].get('ms_bin_num'), d['config'].get('phase_ms'))" 2>/dev/null) $(grep -v gwin $O/$name.err | tail -1 | cut -c1-200)"
}
run r22 X=1 -- --workload rmat --scale 22
run r22_noscr IAS_OPT_G_SCR=0 -- --workload rmat --scale 22
run r20 X=1 -- --workload rmat --scale 20
run r18 X=1 -- --workload rmat --scale 18
timeout 1500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -3 $O/bench_default.err
python -c "
import json; d=json.load(open('$O/bench_default.json'))
print('main', d['config']['format'], round(d['ms_per_step'],3), round(d['value'],1), 'frac', round(d['roofline']['frac'],3), 'e2e', d['e2e'])
for k,v in d.get('also',{}).items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms_per_step','value','error','selected_format')}, v.get('cusparse'), v.get('ell_path',{}).get('ms_per_step'), v.get('e2e'), [v.get(x) for x in ('scale18','scale16','scale14') if v.get(x)], v.get('cpu_baseline'))
print('cpu', d['cpu_baseline'])
"
