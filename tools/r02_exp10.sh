#!/bin/bash
O=gpurun_out/r02_exp10
mkdir -p $O
timeout 900 python -m pytest tests/test_formats_gpu.py -m gpu -x -q --tb=short -k "ell" > $O/test_ell.log 2>&1; echo "ell tests rc=$? $(tail -1 $O/test_ell.log)"
run() {  # name, env..., -- args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu --no-also --no-cusparse --no-e2e --steps 3 --warmup 3 "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$O/$name.json')); print(d['config']['format'], round(d['ms_per_step'],3),'ms', round(d['value'],1),'GF', 'step_frac', round(d['roofline']['step_frac'],3), d['config'].get('ms_bin_sym'), d['config'].get('ms_bin_num'))" 2>/dev/null) $(tail -1 $O/$name.err | cut -c1-200)"
}
run uni_ell X=1 -- --workload uniform --format ell --steps 5
run r22 X=1 -- --workload rmat --scale 22 --steps 2
run r22_b2 IAS_OPT_G2_TAKES_B2=1 -- --workload rmat --scale 22 --steps 2
run r20 X=1 -- --workload rmat --scale 20
run r20_b2 IAS_OPT_G2_TAKES_B2=1 -- --workload rmat --scale 20
