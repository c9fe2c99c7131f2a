"""bench.py's own order of legs (selector + DIA leg, pinned copy, pipelined e2e, CSR leg) with per-call times.
argv: nodia | trim | own | nopipe"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench as BM
from ia_spgemm_b200.engine import get_engine
torch.cuda.set_device(0)
eng = get_engine(0)
stream = torch.cuda.current_stream()
if "own" not in sys.argv:
    eng.set_stream(stream.cuda_stream)
if "nopipe" in sys.argv:
    eng.set_option("e2e_pipeline", 0)
B = BM.Bench(eng, torch, None, 0, 1, stream, 6456.2, "probe")
dA = eng.gen_poisson2d(4096, 4096)
rows = dA.dev.row

def csr(n, tag, quiet=False):
    slow = 0
    for i in range(n):
        t0 = time.perf_counter()
        st = eng.CSR_MUL_CSR_DEV(dA, dA, download=False)[1]
        w = (time.perf_counter() - t0) * 1e3
        if not quiet or w > 4.0:
            print("%s csr %d: wall %.2f ms, device %.2f (analyze %.2f scan %.2f numeric %.2f)" %
                  (tag, i, w, st["ms_total"], st["ms_analyze"], st["ms_scan"], st["ms_numeric"]), flush=True)
        slow += w > 4.0
    return slow

if "nodia" not in sys.argv:
    fmt, dia, feats = B.select(dA)
    m = B.dia_leg(dia, rows, 0, rows, eng.GetFlop(dA, dA), 10, 3, "poisson2d_5pt_4096x4096_A2_fp64")
    print("dia leg", fmt, round(m["ms_per_step"], 3), flush=True)
if "trim" in sys.argv:
    eng.trim_pool()
hA, pins = B.pinned_host_copy(dA)
r = B.e2e_leg(hA, 10, "auto")
print("e2e auto", r["per_call_ms"], r["pipelined"], flush=True)
print("slow among 132 warm:", csr(132, "warm", True))
torch.cuda.synchronize()
csr(4, "timed")
r = B.e2e_leg(hA, 10, "auto")
print("e2e auto again", r["per_call_ms"], r["pipelined"], flush=True)
