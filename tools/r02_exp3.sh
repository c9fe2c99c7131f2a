#!/bin/bash
# round 2, GPU call 3: box facts, the GPU test files one by one (first failure of each shown), A/B of the new kernels
O=gpurun_out/r02_exp3
mkdir -p $O
{ nproc; free -g | head -2; ulimit -l; cat /sys/fs/cgroup/memory.max 2>/dev/null; cat /sys/fs/cgroup/memory/memory.limit_in_bytes 2>/dev/null; python -c "import os;print(len(os.sched_getaffinity(0)))"; } > $O/box.txt 2>&1
cat $O/box.txt | tr '\n' ' '; echo
for f in test_formats_gpu test_csr_gpu test_abi_cxx_gpu test_cli_gpu test_matnet test_fullsize_gpu; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -x -q --tb=short > $O/$f.log 2>&1; echo "$f rc=$? $(tail -1 $O/$f.log)"
done
run() {  # name, env..., -- args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu --no-e2e --no-also --no-cusparse --steps 3 --warmup 3 "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$O/$name.json')); print(d['config']['format'], round(d['ms_per_step'],3),'ms', round(d['value'],1),'GF', 'step_frac', round(d['roofline']['step_frac'],3), d['config'].get('ms_bin_sym'), d['config'].get('ms_bin_num'))" 2>/dev/null) $(tail -1 $O/$name.err | cut -c1-200)"
}
for s in 20 22; do
  run v2_r$s X=1 -- --workload rmat --scale $s
  run v2nolpt_r$s IAS_OPT_G_LPT=0 -- --workload rmat --scale $s
  run v1lpt_r$s IAS_OPT_G_V2=0 -- --workload rmat --scale $s
  run v2notbl_r$s IAS_OPT_G_TBL=0 -- --workload rmat --scale $s
done
run v2_r18 X=1 -- --workload rmat --scale 18
run uni_csr X=1 -- --workload uniform --format csr
run uni_ell X=1 -- --workload uniform --format ell
run uni_ell_pipe IAS_OPT_ELL_ONEPASS=0 -- --workload uniform --format ell
run poi_csr_bulk X=1 -- --workload poisson --format csr --steps 10
run poi_csr_nobulk IAS_OPT_BULK_STORE=0 -- --workload poisson --format csr --steps 10
run poi_auto X=1 -- --workload poisson --steps 10
ls $O | wc -l
