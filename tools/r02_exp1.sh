#!/bin/bash
# round 2, GPU call 1: the GPU test-suite after the boundary changes, then the global-row kernel's phase clocks
# and three cheap variants (L1-cached cell lookups, two 512-thread CTAs per SM, the windowed kernel everywhere).
O=gpurun_out/r02_exp1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests_gpu.log 2>&1; echo "pytest rc=$?" >> $O/tests_gpu.log
tail -5 $O/tests_gpu.log
run() {  # name, env..., -- args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$O/$name.json')); print(round(d['ms_per_step'],2),'ms', round(d['value'],1),'GF', d['config']['ms_bin_sym'], d['config']['ms_bin_num'])" 2>/dev/null)"
}
for s in 20 22; do
  run base_r$s X=1 -- --workload rmat --scale $s
  run ldca_r$s IAS_OPT_G_LDCA=1 -- --workload rmat --scale $s
  run b512_r$s IAS_OPT_G_BLOCK=512 -- --workload rmat --scale $s
  run b512ldca_r$s IAS_OPT_G_BLOCK=512 IAS_OPT_G_LDCA=1 -- --workload rmat --scale $s
done
run gwinall_r20 IAS_OPT_GWIN_MAX_SW=0 -- --workload rmat --scale 20
run gwinall16k_r22 IAS_OPT_GWIN_MAX_SW=0 IAS_OPT_GWIN_SWORDS=16384 -- --workload rmat --scale 22
# phase clocks (instrumented build)
for s in 20 22; do
  run prof_r$s IAS_LIB=$PWD/ia_spgemm_b200/libiaspgemm_prof.so -- --workload rmat --scale $s
  grep "^\[gwin" $O/prof_r$s.err | tail -60 > $O/prof_r$s.phases
done
run profgwin_r20 IAS_LIB=$PWD/ia_spgemm_b200/libiaspgemm_prof.so IAS_OPT_GWIN_MAX_SW=0 -- --workload rmat --scale 20
grep "^\[gwin" $O/profgwin_r20.err | tail -40 > $O/profgwin_r20.phases
ls -la $O
