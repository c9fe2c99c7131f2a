"""ias_row_share on the device against multigpu.snake_row_share (numpy) on operands with many ties."""
import sys
import numpy as np
sys.path.insert(0, ".")
from ia_spgemm_b200 import multigpu as M, workloads as W
from ia_spgemm_b200.engine import get_engine
eng = get_engine(0)
ok = True
for name, A in (("rmat12", W.rmat(12, 8, seed=2)), ("poisson40", W.poisson2d(40)), ("uniform", W.uniform_rows(5000, 6, seed=3))):
    dA = eng.upload(*A)
    ub = M.per_row_products(A[2], A[3], A[2])
    for parts in (1, 2, 3, 8):
        for p in range(parts):
            got = eng.row_share(dA, dA, parts, p).cpu().numpy()
            want = M.snake_row_share(ub, parts, p)
            same = np.array_equal(got, want)
            ok &= same
            if not same:
                print("MISMATCH", name, parts, p, got[:10], want[:10])
    dA.close()
print("row_share == snake_row_share:", ok)
