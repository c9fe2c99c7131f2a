#!/bin/bash
# round 2, GPU call 5: GPU suite after the emit / consume rewrites, phase clocks, e2e host-time stamps
O=gpurun_out/r02_exp5
mkdir -p $O
for f in test_csr_gpu test_formats_gpu test_fullsize_gpu; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -x -q --tb=short > $O/$f.log 2>&1; echo "$f rc=$? $(tail -1 $O/$f.log)"
done
run() {  # name, env..., -- args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu --no-also --no-cusparse --steps 2 --warmup 3 "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$O/$name.json')); print(d['config']['format'], round(d['ms_per_step'],3),'ms', round(d['value'],1),'GF', 'step_frac', round(d['roofline']['step_frac'],3), d['config'].get('ms_bin_sym'), d['config'].get('ms_bin_num'), d['config'].get('phase_ms'), d.get('e2e'))" 2>/dev/null) $(grep -v gwin $O/$name.err | tail -1 | cut -c1-200)"
}
run r20 X=1 -- --workload rmat --scale 20 --no-e2e
run r22 X=1 -- --workload rmat --scale 22 --no-e2e
run prof_r22 IAS_LIB=$PWD/ia_spgemm_b200/libiaspgemm_prof.so -- --workload rmat --scale 22 --no-e2e
run r18 X=1 -- --workload rmat --scale 18 --no-e2e
run poi_auto X=1 -- --workload poisson --steps 10
run poi_csr X=1 -- --workload poisson --format csr --steps 10
