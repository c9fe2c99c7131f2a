#!/bin/bash
O=gpurun_out/r02_exp6
mkdir -p $O
timeout 600 python tools/e2e_probe.py > $O/e2e_probe.log 2>&1; echo "probe rc=$?"; cat $O/e2e_probe.log | tail -20
run() {  # name, env..., -- args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu --no-also --no-cusparse --no-e2e --steps 2 --warmup 3 "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$O/$name.json')); print(d['config']['format'], round(d['ms_per_step'],3),'ms', round(d['value'],1),'GF', 'step_frac', round(d['roofline']['step_frac'],3), d['config'].get('ms_bin_sym'), d['config'].get('ms_bin_num'), d['config'].get('phase_ms'))" 2>/dev/null) $(grep -v gwin $O/$name.err | tail -1 | cut -c1-200)"
}
run r22 X=1 -- --workload rmat --scale 22
run prof_r22 IAS_LIB=$PWD/ia_spgemm_b200/libiaspgemm_prof.so -- --workload rmat --scale 22
run r20 X=1 -- --workload rmat --scale 20
