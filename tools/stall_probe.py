"""Where does the one slow step of bench.py's csr_path leg come from?  Replays the bench's sequence (DIA leg, pipelined
e2e leg, CSR leg) and prints wall + engine phases of every CSR multiply around the synchronisation points."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, ".")
import bench as BM
from ia_spgemm_b200.engine import get_engine
eng = get_engine(0)
stream = torch.cuda.current_stream()
if "own" not in sys.argv:
    eng.set_stream(stream.cuda_stream)          # as bench.py does: the legacy default stream
B = BM.Bench(eng, torch, None, 0, 1, stream, 6456.2, "probe")
dA = eng.gen_poisson2d(4096, 4096)
mode = "noe2e" if "noe2e" in sys.argv else "all"

def csr(n, tag, quiet=False):
    slow = 0
    for i in range(n):
        t0 = time.perf_counter()
        st = eng.CSR_MUL_CSR_DEV(dA, dA, download=False)[1]
        w = (time.perf_counter() - t0) * 1e3
        if not quiet or w > 4.0:
            print("%s csr %d: wall %.2f ms, device %.2f (analyze %.2f symbolic %.2f scan %.2f numeric %.2f)" %
                  (tag, i, w, st["ms_total"], st["ms_analyze"], st["ms_symbolic"], st["ms_scan"], st["ms_numeric"]), flush=True)
        slow += w > 4.0
    return slow

if "dia" in sys.argv:                      # what bench.py does before its e2e leg: selector + DIA leg
    fmt, dia, feats = B.select(dA)
    rows = dA.dev.row
    m = B.dia_leg(dia, rows, 0, rows, eng.GetFlop(dA, dA), 10, 3, "poisson2d_5pt_4096x4096_A2_fp64")
    print("dia leg", fmt, round(m["ms_per_step"], 3), flush=True)
csr(3, "fresh")
print("quiet warm: slow steps", csr(150, "warm0", True))
torch.cuda.synchronize(); csr(3, "after device sync")
ev = torch.cuda.Event(enable_timing=True); ev.record(stream); csr(3, "after event record on engine stream")
time.sleep(0.5); csr(3, "after 0.5 s idle")
if mode != "noe2e":
    hA, pins = B.pinned_host_copy(dA)
    csr(3, "after pinned copy")
    r = B.e2e_leg(hA, 5, "auto")
    print("e2e auto", r["per_call_ms"], r["pipelined"], flush=True)
    csr(3, "right after e2e")
    print("quiet warm: slow steps", csr(150, "warm1", True))
    torch.cuda.synchronize(); csr(3, "after device sync (post e2e)")
    ev = torch.cuda.Event(enable_timing=True); ev.record(stream); csr(3, "after event record (post e2e)")
    time.sleep(0.5); csr(3, "after 0.5 s idle (post e2e)")
    if "short" in sys.argv:
        sys.exit(0)
    for k in range(3):
        print("quiet: slow steps", csr(150, "warm2.%d" % k, True)); torch.cuda.synchronize()
    r = B.e2e_leg(hA, 5, "csr")
    print("e2e csr", r["per_call_ms"], flush=True)
    csr(3, "right after e2e csr")
    print("quiet: slow steps", csr(300, "warm3", True))
    torch.cuda.synchronize(); csr(3, "after device sync (post e2e csr)")
