#!/bin/bash
O=gpurun_out/r02_sampler; mkdir -p $O
for v in "" ""; do
  echo "== variant: $v (host trace)"
  IAS_HOST_TRACE=1 timeout 120 python tools/stall_probe2.py $v 2>&1 | tee -a $O/probe2_trace.log | grep -v "^timed csr [1-3]"
done
