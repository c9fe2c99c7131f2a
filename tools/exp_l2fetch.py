#!/usr/bin/env python
"""Experiment: does cudaLimitMaxL2FetchGranularity = 32 B cut the DRAM over-fetch of the uniform config?
(ncu: k_esc_warp reads 41.7 GB where A + the gathered B rows are 26 GB; B's 64-byte column rows may be fetched
as 128-byte lines.)  usage: exp_l2fetch.py GRANULARITY   -> prints ms per A^2 on uniform 8M x 16"""
import ctypes, sys, time
sys.path.insert(0, ".")
import torch
from ia_spgemm_b200.engine import get_engine
g = int(sys.argv[1])
eng = get_engine()
rt = ctypes.CDLL("libcudart.so.12")
if g:
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(g))      # cudaLimitMaxL2FetchGranularity
    v = ctypes.c_size_t()
    rt.cudaDeviceGetLimit(ctypes.byref(v), 5)
    print("set rc", rc, "granularity now", v.value)
dA = eng.gen_uniform(8000000, 16, 1)
for i in range(6):
    st = eng.CSR_MUL_CSR_DEV(dA, dA, download=False)[1]
    if i >= 3:
        print("ms_total %.2f sym %.2f num %.2f" % (st["ms_total"], st["ms_bin_sym"][2], st["ms_bin_num"][2]))
