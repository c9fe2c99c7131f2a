#!/bin/bash
O=gpurun_out/r02_exp9
mkdir -p $O
timeout 900 python -m pytest tests/test_formats_gpu.py -m gpu -x -q --tb=short > $O/test_formats_gpu.log 2>&1; echo "test_formats rc=$? $(tail -1 $O/test_formats_gpu.log)"; grep -B2 -A12 "Error\|assert" $O/test_formats_gpu.log | head -40
timeout 600 python tools/e2e_probe.py > $O/e2e_probe.log 2>&1; echo "probe rc=$?"; tail -12 $O/e2e_probe.log
IAS_OPT_E2E_PIPELINE=0 timeout 600 python tools/e2e_probe.py > $O/e2e_probe_nopipe.log 2>&1; tail -4 $O/e2e_probe_nopipe.log
timeout 900 python bench.py --no-cpu --no-also --no-cusparse --steps 10 > $O/poi.json 2> $O/poi.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/poi.json')); print(round(d['ms_per_step'],3), d['e2e'])"
