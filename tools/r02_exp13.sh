#!/bin/bash
O=gpurun_out/r02_exp13
mkdir -p $O
timeout 600 python tools/pool_probe.py legacy > $O/pool_probe_legacy.log 2>&1; echo "rc=$?"; cat $O/pool_probe_legacy.log
timeout 600 python tools/pool_probe.py own > $O/pool_probe_own.log 2>&1; echo "rc=$?"; grep -E "csr [03]:|auto" $O/pool_probe_own.log
