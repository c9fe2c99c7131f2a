#!/bin/bash
# round 2: global rows cut into column-range parts -- parity first, then R-MAT 22 (one GPU, and rank 0's share of 8) and rank 0's share of scale 25
O=gpurun_out/r02_split; mkdir -p $O
timeout 600 python -m pytest tests/test_csr_gpu.py -q -x -k "global_rows or rmat or stream or row_list" > $O/tests.log 2>&1; rc=$?; echo "pytest rc=$rc $(tail -1 $O/tests.log)"
if [ $rc -ne 0 ]; then tail -30 $O/tests.log; exit 1; fi
B="python bench.py --workload rmat --no-also --no-cpu --no-e2e --no-cusparse --steps 2 --warmup 3"
for s in 1 0; do
  IAS_OPT_G_SPLIT=$s timeout 300 $B --scale 22 > $O/rmat22_split$s.json 2> $O/rmat22_split$s.err; echo "rmat22 split=$s rc=$? $(python -c "import json;d=json.load(open('$O/rmat22_split$s.json'));print(round(d['ms_per_step'],1), round(d['value'],1))")"
  IAS_OPT_G_SPLIT=$s REPS=2 timeout 200 python tools/r25_share_probe.py 22 8 0 > $O/share22_split$s.log 2>&1; grep "rank 0" $O/share22_split$s.log | cut -c1-160
  IAS_OPT_G_SPLIT=$s REPS=2 timeout 300 python tools/r25_share_probe.py 25 8 0 > $O/share25_split$s.log 2>&1; grep "rank 0" $O/share25_split$s.log | cut -c1-160
done
