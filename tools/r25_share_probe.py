"""What one rank of an 8-GPU strong-scaling run of R-MAT scale 25 does, on ONE GPU: the operand is generated on the
device, ias_row_share deals the rows (sorted by decreasing products, snake order) to 8 ranks, and this process multiplies
the share of the rank(s) named in argv through ias_csr_mul_csr_rowlist_stream.  argv: scale world rank [rank ...]"""
import sys, time
import torch
sys.path.insert(0, ".")
from ia_spgemm_b200.engine import get_engine
scale, world = int(sys.argv[1]), int(sys.argv[2])
ranks = [int(x) for x in sys.argv[3:]] or [0]
torch.cuda.set_device(0)
eng = get_engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
t0 = time.perf_counter()
dA = eng.gen_rmat(scale, 16, seed=1)
eng.sync()
print("generated scale %d: rows %d nnz %d in %.1f s, products %d" % (scale, dA.dev.row, dA.dev.nnz, time.perf_counter() - t0, eng.GetFlop(dA, dA)), flush=True)
for r in ranks:
    share = eng.row_share(dA, dA, world, r)
    n = int(share.numel())
    for rep in range(int(__import__('os').environ.get('REPS', '2'))):
        t1 = time.perf_counter()
        st = eng.csr_mul_csr_rowlist_stream(dA, dA, share.data_ptr(), n)
        w = time.perf_counter() - t1
        print("rank %d/%d rep %d: rows %d products %d nnz(C) %d  wall %.2f s device %.1f ms  batches %s  %.1f GFLOP/s on this GPU" %
              (r, world, rep, n, st["products"], st["nnz"], w, st["ms_total"], st.get("batches"), 2.0 * st["products"] / (st["ms_total"] * 1e6)), flush=True)
        print("   stats", {k: ([round(x, 1) for x in v] if isinstance(v, (list, tuple)) else round(v, 1) if isinstance(v, float) else v)
                           for k, v in st.items()}, flush=True)
