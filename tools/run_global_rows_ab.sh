#!/bin/bash
# A/B of the kernels that take the rows beyond the CTA hash, one GPU.  usage: run_global_rows_ab.sh [SCALE ...]
# Every variant is the same library; IAS_OPT_<NAME> (read by ias_init) picks the kernels:
#   l2      IAS_OPT_GLOBAL_ROWS_SMEM=0   L2 bitmap kernels for both passes (the first version)
#   hybrid  IAS_OPT_GWIN_MAX_SW=1 IAS_OPT_GWIN_SWORDS=4096   shared-memory symbolic + L2 numeric with shared-memory mark
#   gwin    IAS_OPT_GWIN_MAX_SW=0        shared-memory kernels for both passes whatever the column count
#   default                              what the engine picks (gwin numeric only when one super-window covers the columns)
# Add IAS_LIB=$PWD/ia_spgemm_b200/libiaspgemm_prof.so (make -C ia_spgemm_b200/csrc prof) for the phase clocks on stderr.
O=gpurun_out/global_rows_ab
mkdir -p $O
for s in ${@:-18 20}; do
  run() { name=$1; shift; env "$@" timeout 400 python bench.py --workload rmat --scale $s --no-cpu --no-e2e --steps 2 --warmup 3 > $O/rmat${s}_$name.json 2> $O/rmat${s}_$name.err; }
  run l2 IAS_OPT_GLOBAL_ROWS_SMEM=0
  run hybrid IAS_OPT_GWIN_MAX_SW=1 IAS_OPT_GWIN_SWORDS=4096
  run gwin IAS_OPT_GWIN_MAX_SW=0
  run default IAS_NOOP=1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/global_rows_ab/*.json')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); c = d['config']
            print(f.split('/')[-1], 'ms', round(d['ms_per_step'], 2), 'GFLOP/s', round(d['value'], 1), 'sym', c['ms_bin_sym'][3:], 'num', c['ms_bin_num'][3:])
PY
