"""Does the pipelined host path leave the device pool in a state that slows later multiplies?"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from ia_spgemm_b200.engine import get_engine
eng = get_engine(0)
if len(sys.argv) > 1 and sys.argv[1] == "legacy":
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
dA = eng.gen_poisson2d(4096, 4096)
rows, nnz = dA.dev.row, dA.dev.nnz
def csr(n, tag):
    for i in range(n):
        t0 = time.perf_counter()
        st = eng.CSR_MUL_CSR_DEV(dA, dA, download=False)[1]
        print("%s csr %d: wall %.2f ms, device %.2f (analyze %.2f scan %.2f numeric %.2f)" % (tag, i, (time.perf_counter() - t0) * 1e3, st["ms_total"], st["ms_analyze"], st["ms_scan"], st["ms_numeric"]), flush=True)
csr(4, "before")
h_rp = torch.empty(rows + 1, dtype=torch.int32).pin_memory()
h_ci = torch.empty(nnz, dtype=torch.int32).pin_memory()
h_v = torch.empty(nnz, dtype=torch.float64).pin_memory()
for t, ptr in ((h_rp, dA.dev.row_ind_dev), (h_ci, dA.dev.col_ind_dev), (h_v, dA.dev.values_dev)):
    eng.copy(t.data_ptr(), ptr, t.numel() * t.element_size(), 1)
hA = (rows, dA.dev.col, h_rp.numpy(), h_ci.numpy(), h_v.numpy())
for pipe in (0, 1):
    eng.set_option("e2e_pipeline", pipe)
    for i in range(3):
        r = eng.spgemm_auto(hA, hA)
        print("auto pipeline=%d call %d: wall %.1f ms pipelined=%s" % (pipe, i, r["ms"]["wall"], r["pipelined"]), flush=True)
    csr(4, "after pipeline=%d" % pipe)
eng.lib.ias_release_host()
csr(3, "after release_host")
eng.trim_pool()
csr(4, "after trim")
