#!/bin/bash
# round 2, GPU call 4: re-run of the two test files that failed on test bugs, phase clocks of the second-generation global-row
# kernel, DIA 128-bit kernel A/B, then the whole default bench line
O=gpurun_out/r02_exp4
mkdir -p $O
for f in test_formats_gpu test_fullsize_gpu; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q --tb=short > $O/$f.log 2>&1; echo "$f rc=$? $(tail -1 $O/$f.log)"
done
run() {  # name, env..., -- args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu --no-e2e --no-also --no-cusparse --steps 2 --warmup 3 "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$O/$name.json')); print(d['config']['format'], round(d['ms_per_step'],3),'ms', round(d['value'],1),'GF', 'step_frac', round(d['roofline']['step_frac'],3), d['config'].get('ms_bin_sym'), d['config'].get('ms_bin_num'))" 2>/dev/null) $(grep -v gwin $O/$name.err | tail -1 | cut -c1-200)"
}
for s in 20 22; do
  run prof_v2_r$s IAS_LIB=$PWD/ia_spgemm_b200/libiaspgemm_prof.so -- --workload rmat --scale $s
  run prof_v2_win20k_r$s IAS_LIB=$PWD/ia_spgemm_b200/libiaspgemm_prof.so IAS_OPT_G_WIN=20480 IAS_OPT_G_TBL=0 -- --workload rmat --scale $s
done
run uni_ell_bulk X=1 -- --workload uniform --format ell --steps 5
run uni_ell_nobulk IAS_OPT_BULK_STORE=0 -- --workload uniform --format ell --steps 5
run dia_vec X=1 -- --workload poisson --steps 20
run dia_scalar IAS_OPT_DIA_VEC=0 -- --workload poisson --steps 20
timeout 1500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -3 $O/bench_default.err
python -c "
import json; d=json.load(open('$O/bench_default.json'))
print('main', d['config']['format'], round(d['ms_per_step'],3), round(d['value'],1), 'frac', round(d['roofline']['frac'],3), 'e2e', d['e2e'])
for k,v in d.get('also',{}).items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms_per_step','value','error','selected_format')}, v.get('cusparse'), v.get('ell_path',{}).get('ms_per_step'), v.get('e2e'), v.get('scale18'), v.get('cpu_baseline'))
print('cpu', d['cpu_baseline'])
"
