#!/bin/bash
O=gpurun_out/r02_exp12
mkdir -p $O
timeout 900 python -m pytest tests/test_csr_gpu.py -m gpu -x -q --tb=short > $O/test_csr.log 2>&1; echo "csr tests rc=$? $(tail -1 $O/test_csr.log)"
run() {  # name, env..., -- args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu --no-also --no-cusparse --no-e2e --steps 2 --warmup 3 "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$O/$name.json')); print(d['config']['format'], round(d['ms_per_step'],3),'ms', round(d['value'],1),'GF', d['config'].get('ms_bin_sym'), d['config'].get('ms_bin_num'))" 2>/dev/null) $(grep -v gwin $O/$name.err | tail -1 | cut -c1-200)"
}
run r22 X=1 -- --workload rmat --scale 22
run r20 X=1 -- --workload rmat --scale 20
