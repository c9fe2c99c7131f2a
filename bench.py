#!/usr/bin/env python
"""bench.py -- FP64 A^2 SpGEMM throughput (GFLOP/s = 2 x intermediate products / time) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload poisson|uniform|rmat] [--impl reference]

Default workload (N=1): BASELINE.json configs[1], the 2-D 5-point Poisson operator on a 4096x4096 grid
(16.7 M rows), A^2 in FP64 through the CSR hot path.  A "step" is one full multiply: symbolic pass,
allocation of C, numeric pass with column-sorted output (the timed region of the reference's
CUSPARSE_MUL_CUSPARSE, GPU/detail/cusparse/common_cusparse.h:74-93), operands resident in HBM.
N>1: A is row-block partitioned, B (= A) is generated on rank 0 and broadcast once over NCCL before the
timed region, every rank emits its C row block, no further collectives.  For Poisson the grid grows
with N (4096 x 4096*N nodes: per-GPU work fixed -> weak scaling); for R-MAT the graph is fixed and the
rows are split by products (strong scaling).

One JSON line on stdout (rank 0).  `--impl reference` times the reference's CPU path instead
(oracle/_ref = the reference's own MKL_MUL_MKL / CSR_MUL_CSR, else the oracle port) on the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBM_FALLBACK_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


_JSON_FD = None


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip().split(", "))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if val.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------- workloads
def bytes_csr(rows, nnz, rp_bytes=4):
    return rp_bytes * (rows + 1) + 12 * nnz


def make_operand(eng, args, world):
    """Rank-0 generation on device.  Returns DeviceCsr and a description."""
    if args.workload == "poisson":
        nx, ny = args.grid, args.grid * world
        d = eng.gen_poisson2d(nx, ny)
        return d, "2-D 5-point Poisson %dx%d grid (%d rows), A^2 FP64, CSR path" % (nx, ny, nx * ny)
    if args.workload == "uniform":
        d = eng.gen_uniform(args.n, 16, seed=1)
        return d, "uniform random %dx%d, 16 nnz/row, A^2 FP64, CSR path" % (args.n, args.n)
    if args.workload == "rmat":
        d = eng.gen_rmat(args.scale, 16, seed=1)
        return d, "R-MAT scale %d ef 16 (a,b,c=.57,.19,.19, deduped), A^2 FP64, CSR path%s" % (
            args.scale, ", streamed row batches" if args.stream else "")
    raise SystemExit("unknown workload " + args.workload)


def host_operand(args):
    """Same operand on the host (NumPy, bit-identical to the device generator) for the CPU arm."""
    from ia_spgemm_b200 import workloads as W
    if args.workload == "poisson":
        return W.poisson2d(args.grid)
    if args.workload == "uniform":
        return W.uniform_rows(args.n, 16, seed=1)
    return W.rmat(args.scale, 16, seed=1)


def cpu_sample_rows(A, target_products=6e8):
    """A contiguous row block of about target_products intermediate products, starting at rows/3
    (the whole matrix when it is that small)."""
    from ia_spgemm_b200.multigpu import per_row_products
    rows, cols, rp, ci, v = A
    per_row = per_row_products(rp, ci, rp)
    total = int(per_row.sum())
    if total <= target_products * 1.5:
        return 0, rows, total
    start = rows // 3
    cum = np.cumsum(per_row[start:])
    end = start + int(np.searchsorted(cum, target_products)) + 1
    end = min(end, rows)
    return start, end, int(per_row[start:end].sum())


def slice_rows(A, r0, r1):
    rows, cols, rp, ci, v = A
    s, e = int(rp[r0]), int(rp[r1])
    return r1 - r0, cols, (rp[r0:r1 + 1] - rp[r0]).astype(np.int32), ci[s:e], v[s:e]


def cpu_reference_time(A, r0, r1, repeats=2):
    """Times the reference's CPU path on rows [r0,r1) x A.  Returns (best_ms, kind, cores, label)."""
    from oracle.binding import Oracle, Ref, build
    blk = A if (r0 == 0 and r1 == A[0]) else slice_rows(A, r0, r1)
    best, kind, cores, label = None, None, 1, None
    if Ref.available():
        ref = Ref()
        cores = max(ref.threads())
        for _ in range(repeats):
            ms = ref.mkl_mul_mkl(blk, A, keep=False)[3]
            best = ms if best is None else min(best, ms)
        kind, label = "reference", "MKL_MUL_MKL (reference Algorithm 1, %s)" % ref.mkl_version()[:60]
        ms2 = None
        for _ in range(repeats):
            t = ref.csr_mul_csr(blk, A)[3]
            ms2 = t if ms2 is None else min(ms2, t)
        if ms2 < best:
            best, label = ms2, "CSR_MUL_CSR (reference Algorithm 2)"
    else:
        build(ref=False)
        ora = Oracle()
        cores = ora.threads()
        for _ in range(repeats):
            t0 = time.perf_counter()
            ora.csr_mul_csr(blk[0], A[1], blk[2], blk[3], blk[4], A[2], A[3], A[4])
            ms = (time.perf_counter() - t0) * 1e3
            best = ms if best is None else min(best, ms)
        kind, label = "port", "oracle port of CSR_MUL_CSR"
    return best, kind, cores, label


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    A = host_operand(args)
    r0, r1, products = cpu_sample_rows(A)
    from oracle.binding import Ref, Oracle
    times = []
    kind = cores = label = None
    for s in range(args.warmup + args.steps):
        ms, kind, cores, label = cpu_reference_time(A, r0, r1, repeats=1)
        if s >= args.warmup:
            times.append(ms)
    ms = float(np.mean(times))
    val = 2.0 * products / (ms * 1e6)
    sample = "rows [%d,%d) of %d (%d products)%s" % (r0, r1, A[0], products, "" if r1 - r0 < A[0] else " = the full workload")
    line = {"impl": "reference", "metric": "FP64 A^2 SpGEMM GFLOP/s (2*intermediate products/s)", "value": val, "unit": "GFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, 1), "cpu_algorithm": label},
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_name(args, world):
    if args.workload == "poisson":
        return "poisson2d_5pt_%dx%d_A2_fp64" % (args.grid, args.grid * world)
    if args.workload == "uniform":
        return "uniform_%d_x16_A2_fp64" % args.n
    return "rmat_scale%d_ef16_A2_fp64" % args.scale


# ---------------------------------------------------------------------------------------------- engine arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="poisson", choices=["poisson", "uniform", "rmat"])
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--n", type=int, default=8000000)
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--stream", action="store_true", help="streaming row batches (forced for rmat scale >= 19)")
    ap.add_argument("--budget-gb", type=float, default=0.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--format", default="csr", choices=["csr", "dia", "ell"], help="kernel family for the step")
    args = ap.parse_args()
    if args.workload == "rmat" and args.scale >= 19:
        args.stream = True
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup

    # stdout carries exactly one JSON line: libraries that write to fd 1 (NCCL prints its version banner there
    # when NCCL_DEBUG is set) are sent to stderr, the JSON line goes to the saved descriptor
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from ia_spgemm_b200.engine import get_engine
    eng = get_engine(local_rank)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)

    # ---- operands: rank 0 generates, NCCL broadcasts B (= A) once
    from ia_spgemm_b200 import multigpu as M
    t_bcast = 0.0
    desc = None
    if rank == 0:
        dA, desc = make_operand(eng, args, world)
    if world > 1:
        targs = (0, 0, None, None, None)
        if rank == 0:
            t_rp = torch.empty(dA.dev.row + 1, dtype=torch.int32, device="cuda")
            t_ci = torch.empty(dA.dev.nnz, dtype=torch.int32, device="cuda")
            t_v = torch.empty(dA.dev.nnz, dtype=torch.float64, device="cuda")
            for t, ptr in ((t_rp, dA.dev.row_ind_dev), (t_ci, dA.dev.col_ind_dev), (t_v, dA.dev.values_dev)):
                eng.copy(t.data_ptr(), ptr, t.numel() * t.element_size(), 2)
            targs = (dA.dev.row, dA.dev.col, t_rp, t_ci, t_v)
            dA.close()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rows, cols, t_rp, t_ci, t_v = M.broadcast_csr(dist, *targs, src=0, device="cuda")      # NCCL over NVLink
        e1.record()
        torch.cuda.synchronize()
        t_bcast = e0.elapsed_time(e1)
        dA = eng.wrap_device(rows, cols, int(t_ci.numel()), t_rp.data_ptr(), t_ci.data_ptr(), t_v.data_ptr())
        dA._keep = (t_rp, t_ci, t_v)
    rows, cols, nnz_a = dA.dev.row, dA.dev.col, dA.dev.nnz

    # ---- row blocks
    if world > 1:
        bounds = eng.partition_rows(dA, dA, world)
    else:
        bounds = [0, rows]
    r0, r1 = bounds[rank], bounds[rank + 1]

    dia = ell = None
    if args.format == "dia":
        dia = eng.CSRtoDIA(dA, gate=20.0)
        if not dia.choice:
            raise SystemExit("the DIA gate refuses this operand")
    if args.format == "ell":
        ell = eng.CSRtoELL(dA, gate=20.0)
        if not ell.choice:
            raise SystemExit("the ELL gate refuses this operand")

    budget = int(args.budget_gb * 1e9)

    def step():
        if args.format == "dia":
            c, ms = eng.DIA_MUL_DIA_DEV(dia, dia)
            nd = c.num_diagonals
            eng.free_dia(c)
            return {"ms_total": ms, "nnz": 0, "c_diagonals": nd, "products": products_total, "ms_bin_num": [0] * 8, "ms_bin_sym": [0] * 8}
        if args.format == "ell":
            c, ms = eng.ELL_MUL_ELL_DEV(ell, ell)
            st = {"ms_total": ms, "nnz": c.nnz, "c_width": c.max_nnz_per_row, "products": products_total, "ms_bin_num": [0] * 8, "ms_bin_sym": [0] * 8}
            eng.free_ell(c)
            return st
        if args.stream:
            return eng.csr_mul_csr_stream(dA, dA, rows=(r0, r1), budget_bytes=budget)
        return eng.CSR_MUL_CSR_DEV(dA, dA, rows=(r0, r1), download=False)[1]

    products_total = eng.GetFlop(dA, dA) if args.format != "csr" else 0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                     # sampled from the warm-up on: the timed region alone can be a few ms
    t_warm = time.perf_counter()
    n_warm = 0
    while n_warm < args.warmup or (time.perf_counter() - t_warm < 0.3 and n_warm < 200):
        st = step()
        n_warm += 1
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.kernel_launches()
    torch.cuda.synchronize()
    ev0.record(stream)
    stats = []
    for _ in range(args.steps):
        stats.append(step())
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.kernel_launches() - launches0
    ms_total = ev0.elapsed_time(ev1)
    st = stats[-1]
    my_products = st["products"]
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        p = torch.tensor([my_products, st["nnz"], launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(p, op=dist.ReduceOp.SUM)
        products, nnz_c, launches = (int(x) for x in p.tolist())
    else:
        products, nnz_c = my_products, st["nnz"]
    ms_step = ms_total / args.steps
    value = 2.0 * products / (ms_step * 1e6)

    # ---- roofline of the dominant kernel (this rank's numbers; N=1: the whole job)
    peak, peak_src = measured_peak()
    if args.format == "dia":
        nd_a = dia.num_diagonals
        alg_bytes = 8.0 * rows * (2 * nd_a + st["c_diagonals"])
        dom_name, dom_ms = "k_dia_mul_dia", float(np.mean([s["ms_total"] for s in stats]))
    elif args.format == "ell":
        w = ell.max_nnz_per_row
        alg_bytes = (12.0 * w + 4) * rows * 2 + (12.0 * st["c_width"] + 4) * rows
        dom_name, dom_ms = "ell pipeline (k_num_*)", float(np.mean([s["ms_total"] for s in stats]))
    else:
        nrows_blk = r1 - r0
        # bytes_alg(CSR) = bytes(A block) + bytes of the B rows it references at least once + bytes(C block), SURVEY 8(d);
        # this rank's block (N=1: the whole job)
        a_rp = np.zeros(2, dtype=np.int32)
        eng.copy(a_rp.ctypes.data, dA.dev.row_ind_dev + 4 * r0, 4, 1)
        eng.copy(a_rp.ctypes.data + 4, dA.dev.row_ind_dev + 4 * r1, 4, 1)
        a_blk_nnz = int(a_rp[1]) - int(a_rp[0])
        touched_b = eng.touched_b_bytes(dA, dA, rows=(r0, r1))
        alg_bytes = bytes_csr(nrows_blk, a_blk_nnz) + touched_b + bytes_csr(nrows_blk, st["nnz"], 8)
        bins = np.array([[s["ms_bin_num"][b] for b in range(6)] for s in stats]).mean(axis=0)
        names = ["-", "k_num_tiny", "k_esc_warp", "k_num_hash_cta<512,8192>", "k_num_hash_cta<1024,16384>",
                 eng.global_numeric_kernel(cols)]          # generated operands are canonical
        b = int(np.argmax(bins))
        dom_name, dom_ms = names[b], float(bins[b])
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    # dram__bytes_read.sum + dram__bytes_write.sum of that kernel, per launch, from the committed ncu --set full capture
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(workload_name(args, world), {}).get(dom_name.split("<")[0])
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel_ms": dom_ms, "algorithmic_bytes": alg_bytes,
                "step_frac": alg_bytes / (ms_step * 1e-3) / 1e9 / peak}

    # ---- e2e: host operands in pinned memory -> C in pinned host memory, copies inside the timed region
    e2e = None
    if not args.no_e2e and args.format == "csr" and not args.stream:
        h_rp = torch.empty(rows + 1, dtype=torch.int32).pin_memory()
        h_ci = torch.empty(nnz_a, dtype=torch.int32).pin_memory()
        h_v = torch.empty(nnz_a, dtype=torch.float64).pin_memory()
        torch.cuda.synchronize()
        for t, ptr in ((h_rp, dA.dev.row_ind_dev), (h_ci, dA.dev.col_ind_dev), (h_v, dA.dev.values_dev)):
            eng.copy(t.data_ptr(), ptr, t.numel() * t.element_size(), 1)
        hB = (rows, cols, h_rp.numpy(), h_ci.numpy(), h_v.numpy())
        hA = hB if world == 1 else slice_rows(hB, r0, r1)
        if world > 1:
            pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in hA[2:]]
            hA = (hA[0], hA[1], pin[0].numpy(), pin[1].numpy(), pin[2].numpy())
        e_steps = max(2, min(args.steps, 5))
        eng.CSR_MUL_CSR(hA, hB)                     # warm-up: sizes the pinned result arena
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(e_steps):
            (c_rp, c_ci, c_v), est, h2d, d2h = eng.CSR_MUL_CSR(hA, hB)
            sink = float(c_v[-1]) if len(c_v) else 0.0          # read the result on the host
        ev1.record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        e_ms = max(ev0.elapsed_time(ev1), wall) / e_steps
        e_prod = est["products"]
        if world > 1:
            t = torch.tensor([e_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
            p = torch.tensor([e_prod], dtype=torch.int64, device="cuda")
            dist.all_reduce(p, op=dist.ReduceOp.SUM)
            e_prod = int(p.item())
        h2d_b = bytes_csr(hA[0], len(hA[3])) + (0 if world == 1 else bytes_csr(rows, nnz_a))
        d2h_b = bytes_csr(hA[0], est["nnz"], 8)
        e2e = {"value": 2.0 * e_prod / (e_ms * 1e6), "unit": "GFLOP/s", "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
               "ms_per_step": e_ms, "ms_h2d": h2d, "ms_d2h": d2h, "api": "ias_csr_mul_csr_host (CSR_MUL_CSR on host operands)"}
        eng.lib.ias_release_host()

    # ---- the structured-format kernel the front end would select for this operand (reported beside the CSR number)
    also = None
    if world == 1 and args.format == "csr" and args.workload == "poisson":
        try:
            d_dia = eng.CSRtoDIA(dA, gate=20.0)
            if d_dia.choice:
                for _ in range(3):
                    c, _ms = eng.DIA_MUL_DIA_DEV(d_dia, d_dia); eng.free_dia(c)
                torch.cuda.synchronize()
                ev0.record(stream)
                for _ in range(args.steps):
                    c, k_ms = eng.DIA_MUL_DIA_DEV(d_dia, d_dia); nd_c = c.num_diagonals; eng.free_dia(c)
                ev1.record(stream)
                torch.cuda.synchronize()
                dia_ms = ev0.elapsed_time(ev1) / args.steps
                dia_bytes = 8.0 * rows * (2 * d_dia.num_diagonals + nd_c)
                also = {"dia_path": {"ms_per_step": dia_ms, "value": 2.0 * products / (dia_ms * 1e6), "unit": "GFLOP/s",
                                     "kernel": "k_dia_mul_dia", "algorithmic_bytes": dia_bytes,
                                     "frac_of_hbm_peak": dia_bytes / (dia_ms * 1e-3) / 1e9 / peak,
                                     "note": "DIA x DIA on the same operand (what spgemm-gpu's selector runs for banded inputs)"}}
            eng.free_dia(d_dia)
        except Exception as ex:
            also = {"dia_path": {"error": repr(ex)}}

    # ---- CPU baseline: the reference's own CPU path on the host cores (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            A = host_operand(args)
            s0, s1, sp = cpu_sample_rows(A)
            ms, kind, cores, label = cpu_reference_time(A, s0, s1)
            cpu = {"value": 2.0 * sp / (ms * 1e6), "unit": "GFLOP/s", "cores": cores, "kind": kind, "ms": ms, "algorithm": label,
                   "sample": "rows [%d,%d) of %d (%d products)%s, best of 2" % (s0, s1, A[0], sp, "" if s1 - s0 < A[0] else " = the full workload")}
        except Exception as ex:     # the baseline is a reported figure, never a reason to lose the GPU line
            cpu = {"value": None, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (ex,)}

    if rank == 0:
        line = {"metric": "FP64 A^2 SpGEMM GFLOP/s (2*intermediate products/s)", "value": value, "unit": "GFLOP/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong" if (args.workload != "poisson" and world > 1) else "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(args, world), "description": desc, "format": args.format, "rows": rows, "nnz_A": nnz_a,
                           "products": products, "nnz_C": nnz_c, "l2": "inputs_exceed_l2 (A %.2f GB, C %.2f GB per step vs 126 MB L2)" % (
                               bytes_csr(rows, nnz_a) / 1e9, bytes_csr(r1 - r0, st["nnz"], 8) / 1e9),
                           "parallelism": "row-block x%d, B broadcast once (%.1f ms, outside the timed region)" % (world, t_bcast) if world > 1 else "single GPU",
                           "streaming_batches": st.get("batches", 1), "warmup_steps_run": n_warm,
                           "phase_ms": {k: float(np.mean([s.get(k, 0.0) for s in stats])) for k in ("ms_analyze", "ms_symbolic", "ms_scan", "ms_numeric")},
                           "bins": ["empty", "tiny", "warp", "cta_s", "cta_l", "global"],
                           "num_bin_rows": st.get("num_bin_rows"), "sym_bin_rows": st.get("sym_bin_rows"),
                           "ms_bin_sym": [round(float(np.mean([s.get("ms_bin_sym", [0] * 6)[b] for s in stats])), 4) for b in range(6)],
                           "ms_bin_num": [round(float(np.mean([s.get("ms_bin_num", [0] * 6)[b] for s in stats])), 4) for b in range(6)]},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        if also:
            line["also"] = also
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
