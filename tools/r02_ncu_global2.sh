#!/bin/bash
# round 2, final code: --set full of k_num_global2 (with cut rows) at R-MAT scale 20; the same command runs plain first
O=gpurun_out/r02_ncu2; mkdir -p $O
B="python bench.py --no-also --no-cpu --no-e2e --no-cusparse --steps 2 --warmup 3 --workload rmat --scale 20"
timeout 120 $B > $O/plain_rmat20.json 2> $O/plain_rmat20.err && echo "plain rc=0 $(cut -c1-120 $O/plain_rmat20.json)" &&
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_num_global2 -s 1 -c 1 -o $O/global2 $B > $O/ncu_global2.log 2>&1; echo "global2 rc=$?"
ls -la $O
