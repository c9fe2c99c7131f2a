"""Where does the wall time of the auto e2e call go?  Times the raw C call and the Python wrapper separately."""
import ctypes as C, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from ia_spgemm_b200.engine import get_engine, AutoResult
eng = get_engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
dA = eng.gen_poisson2d(4096, 4096)
rows, nnz = dA.dev.row, dA.dev.nnz
h_rp = torch.empty(rows + 1, dtype=torch.int32).pin_memory()
h_ci = torch.empty(nnz, dtype=torch.int32).pin_memory()
h_v = torch.empty(nnz, dtype=torch.float64).pin_memory()
for t, ptr in ((h_rp, dA.dev.row_ind_dev), (h_ci, dA.dev.col_ind_dev), (h_v, dA.dev.values_dev)):
    eng.copy(t.data_ptr(), ptr, t.numel() * t.element_size(), 1)
hA = (rows, dA.dev.col, h_rp.numpy(), h_ci.numpy(), h_v.numpy())
dA.close()
h = eng.host_csr(*hA)
for i in range(6):
    r = AutoResult()
    t0 = time.perf_counter()
    rc = eng.lib.ias_spgemm_auto_host(C.byref(h), C.byref(h), 20.0, None, C.byref(r))
    t1 = time.perf_counter()
    print("raw call %d: rc=%d python wall %.1f ms, C wall %.1f ms, stamps %s" % (i, rc, (t1 - t0) * 1e3, r.ms_wall, [round(x, 1) for x in r.ms_host]), flush=True)
for i in range(4):
    t0 = time.perf_counter()
    res = eng.spgemm_auto(hA, hA)
    t1 = time.perf_counter()
    s = float(res["values"].reshape(-1)[-1])
    t2 = time.perf_counter()
    print("wrapper %d: %.1f ms (+ read %.3f ms), C wall %.1f" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, res["ms"]["wall"]), flush=True)
    del res
    t3 = time.perf_counter()
    print("   del: %.1f ms" % ((t3 - t2) * 1e3), flush=True)
