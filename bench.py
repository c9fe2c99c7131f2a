#!/usr/bin/env python
"""bench.py -- FP64 A^2 SpGEMM throughput (GFLOP/s = 2 x intermediate products / time) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload poisson|uniform|rmat] [--format auto|csr|dia|ell]
                    [--impl engine|reference|cusparse]

Main line (every N): BASELINE.json configs[1], the 2-D 5-point Poisson operator on a 4096 x 4096*N grid, A^2 in FP64
through the front end's own path: feature extraction -> format selection (the engine's rule; MatNet when weight
files are at hand) -> the selected format's kernel.  For this banded operand that is DIA x DIA; the CSR hot path on
the same operand is reported beside it (`also.csr_path`, with its own roofline and e2e).  A "step" is one full
multiply with the operands resident in HBM in the format the kernel takes -- the reference's `run_time`
(CPU/main.cpp:746-748, GPU/detail/cusparse/common_cusparse.h:74-93: symbolic, allocation of C, numeric, sort);
the conversion is its `trans_time` and is reported, and counted inside `e2e`.
N>1: one process per GPU, B (= A) generated on rank 0 and broadcast once over NCCL before the timed region, every rank
multiplies its row block, no further collectives (Poisson grows with N: weak scaling).

Beside the main line the default run measures the other single-GPU BASELINE configs (`also.uniform`, `also.rmat22`,
each with step roofline, sampled CPU baseline and a cuSPARSE figure where cuSPARSE can hold the problem) and, for N>1,
the fixed-size R-MAT scale-22 problem split over the N GPUs (`also.rmat22_strong`: strong scaling, with the 1-GPU time
of the same problem measured in the same run on rank 0).

`--impl reference` times the reference's CPU path (oracle/_ref = the reference's own MKL_MUL_MKL / CSR_MUL_CSR, else
the oracle port) on a bounded sample of the same workload with all host cores; `--impl cusparse` times
cusparseSpGEMM (through torch.sparse.mm) on the same operand -- the same-box GPU library bar.
One JSON line on stdout (rank 0).
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
METRIC = "FP64 A^2 SpGEMM GFLOP/s (2*intermediate products/s)"
BINS = ["empty", "tiny", "warp", "cta_s", "cta_l", "global"]

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


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        self.mode = os.environ.get("IAS_BENCH_CLOCKS", "nvml")
        if self.mode == "off":
            return
        if self.mode == "nvml" and self._start_nvml():
            return
        self.mode = "smi"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip().split(", ")))

    def _start_nvml(self):
        """The same fields read in-process through NVML every 100 ms (what `nvidia-smi --query-gpu -lms 100` prints):
        no child process attaching to the driver ten times a second beside the timed calls."""
        try:
            import pynvml as nv
            import torch
            nv.nvmlInit()
            try:
                h = nv.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(self.index).uuid))
            except Exception:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            bits = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            return False
        self.proc, self.paused, self.quit = "nvml", False, False

        def poll():
            while not self.quit:
                if not self.paused:
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        row = [str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)),
                               "%.1f" % (nv.nvmlDeviceGetPowerUsage(h) / 1000.0)] + ["Active" if r & b else "Not Active" for _, b in bits]
                        self.rows.append((time.perf_counter(), row))
                    except Exception:
                        pass
                time.sleep(0.1)
        self.thread = threading.Thread(target=poll, daemon=True)
        self.thread.start()
        return True

    def mark(self):
        return time.perf_counter()

    def pause(self):
        """Stop polling while host-heavy legs run (e2e, CPU baseline): frequent NVML queries serialise with the many
        small driver calls of those paths; the clocks record is about the device-timed region."""
        if self.proc == "nvml":
            self.paused = True
        elif self.proc:
            try:
                self.proc.send_signal(19)       # SIGSTOP
            except Exception:
                pass

    def resume(self):
        if self.proc == "nvml":
            self.paused = False
        elif self.proc:
            try:
                self.proc.send_signal(18)       # SIGCONT
            except Exception:
                pass

    def summary(self, t0=None, t1=None):
        """Clocks seen between two marks (the whole run when omitted)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable or off"]}
        sm, mx, reasons = [], [], set()
        for t, r in list(self.rows):
            if (t0 is not None and t < t0) or (t1 is not None and t > t1):
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if val.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons),
                "source": "NVML in-process, 100 ms" if self.proc == "nvml" else "nvidia-smi --query-gpu -lms 100"}

    def stop(self):
        if self.proc == "nvml":
            self.quit = True
            self.thread.join(timeout=2)
        elif self.proc:
            self.resume()
            self.proc.terminate()
            self.thread.join(timeout=2)


# ---------------------------------------------------------------------------------------------- workloads
def bytes_csr(rows, nnz, rp_bytes=4):
    return rp_bytes * (rows + 1) + 12 * nnz


def workload_name(kind, grid=4096, world=1, n=8000000, scale=22):
    if kind == "poisson":
        return "poisson2d_5pt_%dx%d_A2_fp64" % (grid, grid * world)
    if kind == "uniform":
        return "uniform_%d_x16_A2_fp64" % n
    return "rmat_scale%d_ef16_A2_fp64" % scale


def describe(kind, grid=4096, world=1, n=8000000, scale=22):
    if kind == "poisson":
        return "2-D 5-point Poisson %dx%d grid (%d rows), A^2 FP64" % (grid, grid * world, grid * grid * world)
    if kind == "uniform":
        return "uniform random %dx%d, 16 nnz/row, A^2 FP64" % (n, n)
    return "R-MAT scale %d ef 16 (a,b,c=.57,.19,.19, deduped), A^2 FP64" % scale


def make_operand(eng, kind, grid=4096, world=1, n=8000000, scale=22):
    if kind == "poisson":
        return eng.gen_poisson2d(grid, grid * world)
    if kind == "uniform":
        return eng.gen_uniform(n, 16, seed=1)
    return eng.gen_rmat(scale, 16, seed=1)


def poisson_rows(nx, ny, r0, r1):
    """Rows [r0, r1) of the nx x ny 5-point operator as (lengths, columns, values) with global column indices."""
    r = np.arange(r0, r1, dtype=np.int64)
    x, y = r % nx, r // nx
    cand = np.stack([r - nx, r - 1, r, r + 1, r + nx], axis=1)
    mask = np.stack([y > 0, x > 0, np.ones(len(r), bool), x < nx - 1, y < ny - 1], axis=1)
    vals = np.broadcast_to(np.array([-1.0, -1.0, 4.0, -1.0, -1.0]), (len(r), 5))
    return mask.sum(axis=1).astype(np.int64), cand[mask].astype(np.int32), np.ascontiguousarray(vals[mask])


def cpu_sample(kind, grid=4096, world=1, n=8000000, scale=22, target_products=6e8, host=None):
    """A bounded sample of the workload for the CPU arm: a contiguous row block of A with about target_products
    intermediate products (the whole matrix when it is that small) and the operand B it multiplies.
    `host`: the operand already on the host (downloaded from the device generator, which is bit-identical to
    workloads.py) -- saves minutes of NumPy generation for the large configs.
    Returns (A_block, B, products, description)."""
    from ia_spgemm_b200 import workloads as W
    from ia_spgemm_b200.multigpu import per_row_products
    if kind == "poisson":
        nx, ny = grid, grid * world
        rows = nx * ny
        total = W.poisson_counts(nx)[1] if world == 1 else None
        if world == 1 and total <= 1.5 * target_products:
            A = host if host is not None else W.poisson2d(nx)
            return A, A, total, "the full workload (%d rows, %d products)" % (rows, total)
        # the N x grid: a block of rows in the middle and exactly the B rows it touches (the rest of B stays empty)
        nblk = int(target_products // 25)
        r0 = (rows // 3) // nx * nx
        r1 = min(rows, r0 + nblk)
        la, ca, va = poisson_rows(nx, ny, r0, r1)
        b0, b1 = max(0, r0 - nx), min(rows, r1 + nx)
        lb, cb, vb = poisson_rows(nx, ny, b0, b1)
        rp_b = np.zeros(rows + 1, dtype=np.int64)
        rp_b[b0 + 1:b1 + 1] = np.cumsum(lb)
        rp_b[b1 + 1:] = rp_b[b1]
        rp_a = np.concatenate(([0], np.cumsum(la)))
        A = (r1 - r0, rows, rp_a.astype(np.int32), ca, va)
        B = (rows, rows, rp_b.astype(np.int32), cb, vb)
        products = int(per_row_products(A[2], A[3], B[2]).sum())
        return A, B, products, "rows [%d,%d) of %d (%d products) x the B rows they touch" % (r0, r1, rows, products)
    A = host if host is not None else (W.uniform_rows(n, 16, seed=1) if kind == "uniform" else W.rmat(scale, 16, seed=1))
    rows, cols, rp, ci, v = A
    per_row = per_row_products(rp, ci, rp)
    total = int(per_row.sum())
    if total <= target_products * 1.5:
        return A, A, total, "the full workload (%d rows, %d products)" % (rows, total)
    start = rows // 3
    cum = np.cumsum(per_row[start:])
    end = min(rows, start + int(np.searchsorted(cum, target_products)) + 1)
    s, e = int(rp[start]), int(rp[end])
    blk = (end - start, cols, (rp[start:end + 1] - rp[start]).astype(np.int32), ci[s:e], v[s:e])
    products = int(per_row[start:end].sum())
    return blk, A, products, "rows [%d,%d) of %d (%d of %d products)" % (start, end, rows, products, total)


def cpu_reference_time(blk, B, repeats=2):
    """Times the reference's CPU path on blk x B with every host core.  Returns (best_ms, kind, cores, label)."""
    from oracle.binding import Oracle, Ref, build
    cores = host_cores()
    best, kind, label = None, None, None
    if Ref.available():
        ref = Ref()
        ref.set_threads(cores)                     # launchers export OMP_NUM_THREADS=1
        cores = max(ref.threads())
        for _ in range(repeats):
            ms = ref.mkl_mul_mkl(blk, B, keep=False)[3]
            best = ms if best is None else min(best, ms)
        kind, label = "reference", "MKL_MUL_MKL (reference Algorithm 1, %s)" % ref.mkl_version()[:60]
        ms2 = None
        for _ in range(repeats):
            t = ref.csr_mul_csr(blk, B)[3]
            ms2 = t if ms2 is None else min(ms2, t)
        if ms2 < best:
            best, label = ms2, "CSR_MUL_CSR (reference Algorithm 2)"
    else:
        build(ref=False)
        ora = Oracle()
        ora.set_threads(cores)
        cores = ora.threads()
        for _ in range(repeats):
            t0 = time.perf_counter()
            ora.csr_mul_csr(blk[0], B[1], blk[2], blk[3], blk[4], B[2], B[3], B[4])
            ms = (time.perf_counter() - t0) * 1e3
            best = ms if best is None else min(best, ms)
        kind, label = "port", "oracle port of CSR_MUL_CSR"
    return best, kind, cores, label


def cpu_baseline(kind, repeats=2, **kw):
    try:
        blk, B, products, sample = cpu_sample(kind, **kw)
        ms, k, cores, label = cpu_reference_time(blk, B, repeats)
        return {"value": 2.0 * products / (ms * 1e6), "unit": "GFLOP/s", "cores": cores, "kind": k, "ms": ms, "algorithm": label,
                "sample": sample + ", best of %d" % repeats}
    except Exception as ex:     # the baseline is a reported figure, never a reason to lose the GPU line
        return {"value": None, "unit": "GFLOP/s", "cores": host_cores(), "kind": "port", "sample": "failed: %r" % (ex,)}


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    blk, B, products, sample = cpu_sample(args.workload, grid=args.grid, world=world, n=args.n, scale=args.scale)
    times = []
    kind = cores = label = None
    for s in range(args.warmup + args.steps):
        ms, kind, cores, label = cpu_reference_time(blk, B, repeats=1)
        if s >= args.warmup:
            times.append(ms)
    ms = float(np.mean(times))
    val = 2.0 * products / (ms * 1e6)
    emit({"impl": "reference", "metric": METRIC, "value": val, "unit": "GFLOP/s",
          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
          "scaling": "n/a (host cores; the sample does not grow with N)", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": {"workload": workload_name(args.workload, args.grid, world, args.n, args.scale), "cpu_algorithm": label},
          "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": sample},
          "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0})


# ---------------------------------------------------------------------------------------------- engine legs
class Bench:
    def __init__(self, eng, torch, dist, rank, world, stream, peak, peak_src):
        self.eng, self.torch, self.dist = eng, torch, dist
        self.rank, self.world, self.stream = rank, world, stream
        self.peak, self.peak_src = peak, peak_src
        self.traffic_table = {}
        try:
            self.traffic_table = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass

    # -- timing ---------------------------------------------------------------------------------------
    def timed(self, step, steps, warmup, min_warm_s=0.3, max_warm=200):
        """W untimed steps, then exactly `steps` steps between CUDA events on the engine's stream, barrier +
        synchronize on both sides, max over ranks.  Returns (ms per step, list of per-step stats, launches, warm-ups run)."""
        torch, dist = self.torch, self.dist
        # warm-up: at least `warmup` steps and min_warm_s, and until two consecutive steps take the same time within 10 %
        # (the first multiplies of a problem size grow the device pool; a step's C call returns synchronised)
        t_warm, n_warm, prev, stable = time.perf_counter(), 0, None, False
        while n_warm < warmup or ((time.perf_counter() - t_warm < min_warm_s or not stable) and n_warm < max_warm):
            t1 = time.perf_counter()
            step()
            dt = time.perf_counter() - t1
            stable = prev is not None and abs(dt - prev) <= 0.1 * max(dt, prev)
            prev = dt
            n_warm += 1
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = self.eng.kernel_launches()
        torch.cuda.synchronize()
        ev0.record(self.stream)
        stats = [step() for _ in range(steps)]
        ev1.record(self.stream)
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        ms = ev0.elapsed_time(ev1)
        launches = self.eng.kernel_launches() - l0
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, stats, launches, n_warm

    def all_sum(self, *ints):
        if self.world == 1:
            return [int(x) for x in ints]
        t = self.torch.tensor(list(ints), dtype=self.torch.int64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(x) for x in t.tolist()]

    def traffic(self, wname, kernel):
        """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture
        of this workload (profiles/traffic.json, written by tools/summarize_ncu.py from profiles/*.csv)."""
        return self.traffic_table.get(wname, {}).get(kernel.split("<")[0])

    # -- formats --------------------------------------------------------------------------------------
    def select(self, dA):
        """The front end's selector on the device operand: 26 features -> rule (ias_select_format)."""
        eng = self.eng
        f = eng.features26(dA, dA)
        dia = eng.CSRtoDIA(dA, gate=20.0)
        w = eng.max_row_nnz(dA)
        ell_ok = eng.lib.ias_sizeof_ell(dA.dev.row, w) < 20.0 * eng.lib.ias_sizeof_csr(dA.dev.row, dA.dev.nnz)
        cls = eng.select_format(f, bool(dia.choice), bool(ell_ok))
        if not dia.choice or cls != 2:
            eng.free_dia(dia)
            dia = None
        return {1: "csr", 2: "dia", 3: "ell"}[cls], dia, f

    def csr_leg(self, dA, r0, r1, steps, warmup, stream_mode, budget, wname, blocks=None, rowlist=None):
        """CSR hot path on rows [r0, r1) (or on a list of row blocks, or on a device list of rows): ms/step, totals, bins,
        step and kernel roofline."""
        eng = self.eng

        def one(b0, b1):
            if rowlist is not None:
                return eng.csr_mul_csr_rowlist_stream(dA, dA, rowlist.data_ptr(), int(rowlist.numel()), budget_bytes=budget)
            if stream_mode:
                return eng.csr_mul_csr_stream(dA, dA, rows=(b0, b1), budget_bytes=budget)
            return eng.CSR_MUL_CSR_DEV(dA, dA, rows=(b0, b1), download=False)[1]

        blocks = blocks or [(r0, r1)]

        def step():
            sts = [one(b0, b1) for b0, b1 in blocks if b1 > b0]
            if len(sts) == 1:
                return sts[0]
            agg = {"products": sum(s["products"] for s in sts), "nnz": sum(s["nnz"] for s in sts),
                   "batches": sum(s.get("batches", 1) for s in sts)}
            for k in ("ms_analyze", "ms_symbolic", "ms_scan", "ms_numeric"):
                agg[k] = sum(s.get(k, 0.0) for s in sts)
            for k in ("ms_bin_sym", "ms_bin_num", "sym_bin_rows", "num_bin_rows"):
                agg[k] = [sum(s[k][b] for s in sts) for b in range(6)]
            return agg

        ms_step, stats, launches, n_warm = self.timed(step, steps, warmup)
        st = stats[-1]
        my_rows = sum(b1 - b0 for b0, b1 in blocks) if rowlist is None else int(rowlist.numel())
        products, nnz_c, launches_all = self.all_sum(st["products"], st["nnz"], launches)
        # bytes_alg(CSR) = bytes(A block) + bytes of the B rows it references at least once + bytes(C block), SURVEY 8(d);
        # this rank's rows (N=1: the whole job)
        a_nnz, touched = 0, 0
        if rowlist is not None:               # a cyclic share of a skewed operand: 1/N of A's entries, (nearly) every B row touched
            a_nnz = dA.dev.nnz // max(self.world, 1)
            touched = bytes_csr(dA.dev.row, dA.dev.nnz)
        for b0, b1 in (blocks if rowlist is None else []):
            if b1 <= b0:
                continue
            a_rp = np.zeros(2, dtype=np.int32)
            eng.copy(a_rp.ctypes.data, dA.dev.row_ind_dev + 4 * b0, 4, 1)
            eng.copy(a_rp.ctypes.data + 4, dA.dev.row_ind_dev + 4 * b1, 4, 1)
            a_nnz += int(a_rp[1]) - int(a_rp[0])
            touched += eng.touched_b_bytes(dA, dA, rows=(b0, b1))
        alg_bytes = bytes_csr(my_rows, a_nnz) + touched + bytes_csr(my_rows, st["nnz"], 8)
        bins = np.array([[s["ms_bin_num"][b] for b in range(6)] for s in stats]).mean(axis=0)
        names = ["-", "k_num_tiny", "k_esc_warp", "k_num_hash_cta<512,8192>", "k_num_hash_cta<1024,16384>",
                 eng.global_numeric_kernel(dA.dev.col)]          # generated operands are canonical
        b = int(np.argmax(bins))
        dom_name, dom_ms = names[b], float(bins[b])
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        roof = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": self.peak, "unit": "GB/s",
                "frac": achieved / self.peak, "traffic": self.traffic(wname, dom_name), "peak_source": self.peak_src,
                "kernel_ms": dom_ms, "algorithmic_bytes": alg_bytes,
                "step_frac": alg_bytes / (ms_step * 1e-3) / 1e9 / self.peak,
                "note": "rank 0's rows" if self.world > 1 else "the whole job"}
        # secondary diagnostic (not graded): the no-cache gather model -- every B entry re-fetched per product
        gather_bytes = alg_bytes + 12 * int(st["products"])
        roof["gather_model"] = {"bytes": gather_bytes, "step_frac": gather_bytes / (ms_step * 1e-3) / 1e9 / self.peak,
                                "note": "bytes_alg + 12 B per intermediate product"}
        cfg = {"streaming_batches": st.get("batches", 1), "warmup_steps_run": n_warm,
               "engine_ms_per_call": [round(float(x.get("ms_total", 0.0)), 3) for x in stats],      # the C call's own event timing, every step
               "phase_ms": {k: float(np.mean([s.get(k, 0.0) for s in stats])) for k in ("ms_analyze", "ms_symbolic", "ms_scan", "ms_numeric")},
               "bins": BINS, "num_bin_rows": st.get("num_bin_rows"), "sym_bin_rows": st.get("sym_bin_rows"),
               "ms_bin_sym": [round(float(np.mean([s["ms_bin_sym"][b] for s in stats])), 4) for b in range(6)],
               "ms_bin_num": [round(float(np.mean([s["ms_bin_num"][b] for s in stats])), 4) for b in range(6)]}
        return {"ms_per_step": ms_step, "value": 2.0 * products / (ms_step * 1e6), "unit": "GFLOP/s", "products": products,
                "nnz_C": nnz_c, "roofline": roof, "launches": launches_all, "detail": cfg, "my_ms": ms_step}

    def dia_leg(self, dia, rows, r0, r1, products, steps, warmup, wname):
        eng = self.eng

        def step():
            c, ms = eng.DIA_MUL_DIA_DEV(dia, dia, rows=(r0, r1))
            out = {"ms_total": ms, "c_diagonals": c.num_diagonals}
            eng.free_dia(c)
            return out

        ms_step, stats, launches, n_warm = self.timed(step, steps, warmup)
        nd_c = stats[-1]["c_diagonals"]
        kern_ms = float(np.mean([s["ms_total"] for s in stats]))
        alg_bytes = 8.0 * (r1 - r0) * (2 * dia.num_diagonals + nd_c)          # bytes_alg(DIA) = 8 n (dA + dB + dC), this rank's rows
        (launches_all,) = self.all_sum(launches)
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "k_dia_mul_dia", "achieved": achieved, "peak": self.peak, "unit": "GB/s",
                "frac": achieved / self.peak, "traffic": self.traffic(wname, "k_dia_mul_dia"), "peak_source": self.peak_src,
                "kernel_ms": kern_ms, "algorithmic_bytes": alg_bytes, "step_frac": alg_bytes / (ms_step * 1e-3) / 1e9 / self.peak}
        return {"ms_per_step": ms_step, "value": 2.0 * products / (ms_step * 1e6), "unit": "GFLOP/s", "products": products,
                "c_diagonals": nd_c, "roofline": roof, "launches": launches_all,
                "detail": {"warmup_steps_run": n_warm, "engine_ms_per_call": [round(float(x["ms_total"]), 3) for x in stats]}}

    def ell_leg(self, dA, rows, products, steps, warmup, wname):
        eng = self.eng
        t0 = time.perf_counter()
        ell = eng.CSRtoELL(dA, gate=20.0)
        eng.sync()
        trans_ms = (time.perf_counter() - t0) * 1e3
        if not ell.choice:
            return {"error": "the ELL gate refuses this operand"}

        def step():
            c, ms = eng.ELL_MUL_ELL_DEV(ell, ell)
            out = {"ms_total": ms, "nnz": c.nnz, "c_width": c.max_nnz_per_row}
            eng.free_ell(c)
            return out

        ms_step, stats, launches, n_warm = self.timed(step, steps, warmup)
        st = stats[-1]
        w = ell.max_nnz_per_row
        alg_bytes = (12.0 * w + 4) * rows * 2 + (12.0 * st["c_width"] + 4) * rows
        eng.free_ell(ell)
        roof = {"bound": "hbm", "kernel": "k_ell_mul_ell", "achieved": alg_bytes / (ms_step * 1e-3) / 1e9, "peak": self.peak, "unit": "GB/s",
                "frac": alg_bytes / (ms_step * 1e-3) / 1e9 / self.peak, "traffic": self.traffic(wname, "k_ell_mul_ell"), "peak_source": self.peak_src,
                "algorithmic_bytes": alg_bytes, "step_frac": alg_bytes / (ms_step * 1e-3) / 1e9 / self.peak}
        return {"ms_per_step": ms_step, "value": 2.0 * products / (ms_step * 1e6), "unit": "GFLOP/s", "products": products,
                "nnz_C": st["nnz"], "c_width": st["c_width"], "trans_ms": trans_ms, "roofline": roof, "launches": launches,
                "detail": {"warmup_steps_run": n_warm, "engine_ms_per_call": [round(float(x["ms_total"]), 3) for x in stats]}}

    # -- e2e ------------------------------------------------------------------------------------------
    def pinned_host_copy(self, dA):
        torch, eng = self.torch, self.eng
        rows, nnz = dA.dev.row, dA.dev.nnz
        h_rp = torch.empty(rows + 1, dtype=torch.int32).pin_memory()
        h_ci = torch.empty(nnz, dtype=torch.int32).pin_memory()
        h_v = torch.empty(nnz, dtype=torch.float64).pin_memory()
        torch.cuda.synchronize()
        for t, ptr in ((h_rp, dA.dev.row_ind_dev), (h_ci, dA.dev.col_ind_dev), (h_v, dA.dev.values_dev)):
            eng.copy(t.data_ptr(), ptr, t.numel() * t.element_size(), 1)
        hA = (rows, dA.dev.col, h_rp.numpy(), h_ci.numpy(), h_v.numpy())
        return hA, (h_rp, h_ci, h_v)

    def e2e_leg(self, hA, steps, api):
        """Host operands in pinned memory -> result in pinned host memory through one C-ABI call; every copy is inside
        the timed region.  api: "auto" = ias_spgemm_auto_host (the front end's path), "csr" = ias_csr_mul_csr_host."""
        torch, eng = self.torch, self.eng
        e_steps = max(2, min(steps, 5))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        # the C-ABI calls themselves (ctypes, no convenience wrapper): host structs are built once, like a C caller's
        from ia_spgemm_b200.engine import AutoResult, SpgemmStats
        h = eng.host_csr(*hA)
        lib = eng.lib

        def call():
            if api == "auto":
                r = AutoResult()
                eng._ck(lib.ias_spgemm_auto_host(C.byref(h), C.byref(h), C.c_double(20.0), None, C.byref(r)))
                cells = r.nnz if r.format != 3 else r.row * r.max_nnz_per_row
                sink = float(r.values[cells - 1]) if cells else 0.0                       # read the result on the host
                return {"format": {1: "csr", 2: "dia", 3: "ell"}[r.format], "nnz": r.nnz, "h2d_bytes": r.h2d_bytes, "d2h_bytes": r.d2h_bytes,
                        "ms": {k: getattr(r, "ms_" + k) for k in ("h2d", "select", "convert", "multiply", "d2h", "wall")},
                        "host_ms_at": [round(x, 3) for x in r.ms_host], "pipelined": bool(r.pipelined)}, sink
            rp, ci, v = C.POINTER(C.c_longlong)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
            nnz, st, h2d, d2h = C.c_longlong(), SpgemmStats(), C.c_double(), C.c_double()
            eng._ck(lib.ias_csr_mul_csr_host(C.byref(h), C.byref(h), C.byref(rp), C.byref(ci), C.byref(v), C.byref(nnz), C.byref(st),
                                             C.byref(h2d), C.byref(d2h)))
            sink = float(v[nnz.value - 1]) if nnz.value else 0.0
            return {"format": "csr", "nnz": nnz.value, "products": st.products, "ms": {"h2d": h2d.value, "d2h": d2h.value},
                    "h2d_bytes": bytes_csr(hA[0], len(hA[3])), "d2h_bytes": bytes_csr(hA[0], nnz.value, 8)}, sink

        call()                                      # warm-up: sizes the pinned result arena
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev0.record(self.stream)
        per_call = []
        for _ in range(e_steps):
            tc = time.perf_counter()
            r, sink = call()
            per_call.append(round((time.perf_counter() - tc) * 1e3, 2))
        ev1.record(self.stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        e_ms = max(ev0.elapsed_time(ev1), wall) / e_steps
        out = {"ms_per_step": e_ms, "h2d_bytes_per_step": int(r["h2d_bytes"]), "d2h_bytes_per_step": int(r["d2h_bytes"]),
               "result_format": r["format"], "phase_ms": {k: round(float(v), 3) for k, v in r["ms"].items()},
               "host_ms_at": r.get("host_ms_at"), "per_call_ms": per_call, "pipelined": r.get("pipelined"),
               "api": "ias_spgemm_auto_host (features -> selection -> conversion -> multiply -> host result)" if api == "auto"
                      else "ias_csr_mul_csr_host (CSR_MUL_CSR on host operands)"}
        eng.lib.ias_release_host()
        return out

    # -- cuSPARSE -------------------------------------------------------------------------------------
    def cusparse_leg(self, dA, products, steps=2):
        """cusparseSpGEMM (generic API) on the same device operand, through torch.sparse.mm on CSR tensors; the timed
        region is the whole call: buffer sizing, symbolic, allocation of C, numeric -- what CUSPARSE_MUL_CUSPARSE times
        (GPU/detail/cusparse/common_cusparse.h:74-93).  Library code: a yardstick, never the product path."""
        torch, eng = self.torch, self.eng
        try:
            rows, cols, nnz = dA.dev.row, dA.dev.col, dA.dev.nnz
            crow = torch.empty(rows + 1, dtype=torch.int32, device="cuda")
            col = torch.empty(nnz, dtype=torch.int32, device="cuda")
            val = torch.empty(nnz, dtype=torch.float64, device="cuda")
            for t, ptr in ((crow, dA.dev.row_ind_dev), (col, dA.dev.col_ind_dev), (val, dA.dev.values_dev)):
                eng.copy(t.data_ptr(), ptr, t.numel() * t.element_size(), 2)
            A = torch.sparse_csr_tensor(crow, col, val, size=(rows, cols))
            torch.cuda.synchronize()
            Cm = torch.sparse.mm(A, A)              # warm-up
            nnz_c = int(Cm._nnz())
            del Cm
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(steps):
                Cm = torch.sparse.mm(A, A)
                del Cm
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / steps
            del A, crow, col, val
            torch.cuda.empty_cache()
            return {"ms_per_step": ms, "value": 2.0 * products / (ms * 1e6), "unit": "GFLOP/s", "nnz_C": nnz_c,
                    "api": "cusparseSpGEMM via torch.sparse.mm (CSR, fp64, int32 indices), CUDA %s" % torch.version.cuda}
        except Exception as ex:
            try:
                torch.cuda.empty_cache()
            except Exception:
                pass
            return {"error": ("%s: %s" % (type(ex).__name__, ex))[:300]}


# ---------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference", "cusparse"])
    ap.add_argument("--workload", default=None, choices=["poisson", "uniform", "rmat"],
                    help="default: poisson as the main line plus the other BASELINE configs under `also`")
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--n", type=int, default=8000000)
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--stream", action="store_true", help="streaming row batches (forced for rmat scale >= 19)")
    ap.add_argument("--budget-gb", type=float, default=0.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="main line only")
    ap.add_argument("--no-cusparse", action="store_true")
    ap.add_argument("--format", default="auto", choices=["auto", "csr", "dia", "ell"], help="kernel family for the main step")
    args = ap.parse_args()
    default_run = args.workload is None
    if default_run:
        args.workload = "poisson"
    if args.workload == "rmat" and args.scale >= 19:
        args.stream = True
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

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
    from ia_spgemm_b200 import multigpu as M
    eng = get_engine(local_rank)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    if world > 1:
        eng.set_option("trust_operand_cache", 1)     # operands live in tensors this script keeps alive until the end
    peak, peak_src = measured_peak()
    B = Bench(eng, torch, dist, rank, world, stream, peak, peak_src)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    def shared_operand(kind, **kw):
        """Rank 0 generates on device; N>1: NCCL broadcast of the three CSR arrays.  Returns (DeviceCsr, bcast ms)."""
        dA = make_operand(eng, kind, **kw) if rank == 0 else None
        if world == 1:
            return dA, 0.0
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
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        d = eng.wrap_device(rows, cols, int(t_ci.numel()), t_rp.data_ptr(), t_ci.data_ptr(), t_v.data_ptr())
        d._keep = (t_rp, t_ci, t_v)
        return d, float(t.item())

    if args.impl == "cusparse":
        # the same-box GPU library bar as its own arm: cusparseSpGEMM on the same device operand (one GPU)
        if rank == 0:
            kind = args.workload
            kw = dict(grid=args.grid, world=1, n=args.n, scale=args.scale)
            dA = make_operand(eng, kind, **kw)
            products = eng.GetFlop(dA, dA)
            r = B.cusparse_leg(dA, products, steps=max(2, min(args.steps, 5)))
            dA.close()
            sampler.stop()
            emit({"impl": "cusparse", "metric": METRIC, "value": r.get("value"), "unit": "GFLOP/s", "n_gpus": 1, "steps": args.steps,
                  "warmup": args.warmup, "ms_per_step": r.get("ms_per_step"), "higher_is_better": True, "scaling": "n/a (one GPU)",
                  "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                  "config": {"workload": workload_name(kind, **kw), "description": describe(kind, **kw), "api": r.get("api"), "error": r.get("error")},
                  "gpu_launches": 0})
        if world > 1:
            dist.destroy_process_group()
        return

    # ================================================================== main leg
    kind = args.workload
    kw = dict(grid=args.grid, world=world if kind == "poisson" else 1, n=args.n, scale=args.scale)
    wname = workload_name(kind, **kw)
    t_mark0 = sampler.mark()
    dA, t_bcast = shared_operand(kind, **kw)
    rows, cols, nnz_a = dA.dev.row, dA.dev.col, dA.dev.nnz
    budget = int(args.budget_gb * 1e9)
    t_sel = time.perf_counter()
    fmt, dia, feats = (args.format, None, None)
    if args.format == "auto":
        fmt, dia, feats = B.select(dA)
    elif args.format == "dia":
        dia = eng.CSRtoDIA(dA, gate=20.0)
        if not dia.choice:
            raise SystemExit("the DIA gate refuses this operand")
    eng.sync()
    select_ms = (time.perf_counter() - t_sel) * 1e3
    strong = kind != "poisson" and world > 1
    rowlist = None
    if world > 1:
        if strong:
            # rank r takes rows r, r + N, ...: one pass of the pipeline per rank, hubs and tail rows on every rank
            rowlist = eng.row_share(dA, dA, world, rank)
            r0, r1, blocks = 0, rows, None
        else:
            bounds = eng.partition_rows(dA, dA, world)
            r0, r1 = bounds[rank], bounds[rank + 1]
            blocks = None
    else:
        r0, r1, blocks = 0, rows, None

    products_total = eng.GetFlop(dA, dA)
    if fmt == "dia":
        main = B.dia_leg(dia, rows, r0, r1, products_total, args.steps, args.warmup, wname)
    elif fmt == "ell":
        if world > 1:
            raise SystemExit("the ELL path has no row-block entry: use --format csr for N > 1")
        main = B.ell_leg(dA, rows, products_total, args.steps, args.warmup, wname)
        if "error" in main:
            raise SystemExit(main["error"])
    else:
        main = B.csr_leg(dA, r0, r1, args.steps, args.warmup, args.stream, budget, wname, blocks=blocks, rowlist=rowlist)
    t_mark1 = sampler.mark()
    sampler.pause()

    also = {}
    e2e = None
    hA = pins = None
    def safe_e2e(api):
        try:
            r = B.e2e_leg(hA, args.steps, api)
            r["value"] = 2.0 * products_total / (r["ms_per_step"] * 1e6)
            r["unit"] = "GFLOP/s"
            return r
        except Exception as ex:               # a host-memory limit must not cost the device-timed line
            try:
                eng.lib.ias_release_host()
            except Exception:
                pass
            return {"value": None, "unit": "GFLOP/s", "error": ("%s: %s" % (type(ex).__name__, ex))[:300]}

    if not args.no_e2e and world == 1 and not args.stream:
        hA, pins = B.pinned_host_copy(dA)
        e2e = safe_e2e("auto" if args.format == "auto" else "csr")
    elif not args.no_e2e and world > 1 and kind == "poisson":
        # N ranks, host operands: every rank uploads 1/N of B from its pinned host memory, the blocks are all-gathered
        # over NVLink, each rank multiplies its row block and downloads its block of the result
        e2e = multi_gpu_e2e(B, eng, torch, dist, dA, fmt, r0, r1, rank, world, products_total, max(2, min(args.steps, 5)))

    # ---- the CSR hot path on the same operand, when the selector took a structured format
    if fmt != "csr" and not args.no_also:
        try:
            c = B.csr_leg(dA, r0, r1, args.steps, args.warmup, False, budget, wname)
            if hA is not None:
                c["e2e"] = safe_e2e("csr")
            c["note"] = "Algorithm 2 (CSR Gustavson pipeline) on the same operand"
        except Exception as ex:
            c = {"error": ("%s: %s" % (type(ex).__name__, ex))[:300]}
        also["csr_path"] = c
    if rank == 0 and world == 1 and not args.no_cusparse and not args.no_also and not args.stream:
        also["cusparse"] = B.cusparse_leg(dA, products_total)
    if dia is not None:
        eng.free_dia(dia)
    del hA, pins
    host_main = dA.download() if (rank == 0 and world == 1 and not args.no_cpu) else None
    dA.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(kind, host=host_main, **kw)
        del host_main

    # ================================================================== the other BASELINE configs
    sampler.resume()
    if default_run and not args.no_also:
        if world == 1:
            also["uniform"] = side_config(B, eng, "uniform", args, with_cpu=not args.no_cpu, with_cusparse=not args.no_cusparse)
            also["rmat22"] = side_config(B, eng, "rmat", args, with_cpu=not args.no_cpu, with_cusparse=not args.no_cusparse)
        else:
            sampler.pause()              # NVML polling of GPU 0 during a multi-rank timed leg would single out rank 0
            also["rmat22_strong"] = strong_scaling_leg(B, eng, torch, dist, shared_operand, args, rank, world)
            scale5 = int(os.environ.get("IAS_BENCH_STRONG_SCALE", "25"))       # BASELINE config 5: R-MAT scale 25 over the GPUs of the box
            if scale5 > 0:
                also["rmat%d_strong" % scale5] = strong_scaling_leg(B, eng, torch, dist, shared_operand, args, rank, world, scale=scale5, one_gpu=False)
            sampler.resume()

    clocks = None
    if rank == 0:
        clocks = sampler.summary(t_mark0, t_mark1)
        clocks["whole_run"] = sampler.summary()
        sampler.stop()
        line = {"metric": METRIC, "value": main["value"], "unit": "GFLOP/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
                "scaling": "strong" if strong else ("weak" if world > 1 else "n/a (one GPU)"), "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": wname, "description": describe(kind, **kw), "format": fmt,
                           "format_selection": ("rule on the 26 features (ias_select_format): diagonal fill %.3f, ELL efficiency %.3f, CV %.4f"
                                                % (feats[2] / max(feats[18] * feats[0], 1.0), feats[24], feats[8])) if feats is not None else "forced by --format",
                           "select_and_convert_ms": round(select_ms, 3),
                           "rows": rows, "nnz_A": nnz_a, "products": main["products"], "nnz_C": main.get("nnz_C"),
                           "c_diagonals": main.get("c_diagonals"),
                           "l2": "inputs_exceed_l2 (A %.2f GB per step vs 126 MB L2)" % (bytes_csr(rows, nnz_a) / 1e9),
                           "parallelism": ("row blocks x%d (%s), B broadcast once over NCCL (%.1f ms, outside the timed region)"
                                           % (world, "rows by decreasing products dealt in snake order" if strong else "contiguous, balanced by products", t_bcast))
                                          if world > 1 else "single GPU",
                           "tolerance": "structure bit-exact; values within 1e-12 x sum|a*b| of the entry (= 1e-12 x |c| on zero-free, cancellation-free operands)"},
                "roofline": main["roofline"], "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(main["launches"]), "clocks": clocks}
        line["config"].update(main.get("detail", {}))
        if also:
            line["also"] = also
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def side_config(B, eng, kind, args, with_cpu, with_cusparse):
    """One of the other single-GPU BASELINE configs, measured like the main leg (fewer steps for the long ones)."""
    out = {}
    try:
        kw = dict(grid=args.grid, world=1, n=args.n, scale=22)
        wname = workload_name(kind, **kw)
        eng.trim_pool()                                      # every config starts from an empty device pool, like its own process
        dA = make_operand(eng, kind, **kw)
        products = eng.GetFlop(dA, dA)
        rows = dA.dev.row
        fmt, dia, feats = B.select(dA)
        if dia is not None:
            eng.free_dia(dia)
        out = {"workload": wname, "description": describe(kind, **kw), "selected_format": fmt, "rows": rows, "nnz_A": dA.dev.nnz}
        if fmt == "ell":                                                 # the selector's path first
            out["ell_path"] = B.ell_leg(dA, rows, products, 5, 3, wname)
            eng.trim_pool()
        if kind == "rmat":
            leg = B.csr_leg(dA, 0, rows, 2, 1, True, 0, wname)           # 1.5e11 products per step: one warm-up, two timed steps
        else:
            leg = B.csr_leg(dA, 0, rows, 5, 3, False, 0, wname)          # C (24.6 GB) is materialised
        out.update({k: leg[k] for k in ("ms_per_step", "value", "unit", "products", "nnz_C", "roofline")})
        out["detail"] = leg["detail"]
        if fmt == "ell":
            out["csr_path"] = {k: out[k] for k in ("ms_per_step", "value", "unit", "roofline")}
            out.update({k: out["ell_path"][k] for k in ("ms_per_step", "value", "unit", "roofline")})
            out["note"] = "top-level figures: the selector's path (ELL); csr_path: the CSR pipeline on the same operand"
        if with_cusparse and kind != "rmat":
            out["cusparse"] = B.cusparse_leg(dA, products)
        host = dA.download() if with_cpu else None          # the device generator's operand, for the CPU sample below
        dA.close()
        if kind == "rmat" and with_cusparse:
            # cusparseSpGEMM materialises C and its workspace: scale 22 (0.86 TB) is out of reach, scale 18 (15 GB) is
            # the largest it holds; the engine is timed on the same operand beside it
            for sc in (18, 16, 14):
                d18 = make_operand(eng, "rmat", scale=sc)
                p18 = eng.GetFlop(d18, d18)
                e18 = B.csr_leg(d18, 0, d18.dev.row, 3, 3, False, 0, workload_name("rmat", scale=sc))
                cu = B.cusparse_leg(d18, p18)
                out["scale%d" % sc] = {"engine": {k: e18[k] for k in ("ms_per_step", "value", "unit", "products", "nnz_C")}, "cusparse": cu}
                d18.close()
                if "error" not in cu:
                    break
        if with_cpu:
            out["cpu_baseline"] = cpu_baseline(kind, repeats=1, host=host, **kw)
    except Exception as ex:
        out["error"] = ("%s: %s" % (type(ex).__name__, ex))[:300]
    return out


def strong_scaling_leg(B, eng, torch, dist, shared_operand, args, rank, world, scale=22, one_gpu=True):
    """R-MAT scale 22 (fixed problem) over the N GPUs: B broadcast once; the rows, sorted by decreasing products, are dealt
    to the ranks in snake order (ias_row_share) and every rank multiplies its share in ONE pass of the streaming pipeline
    (ias_csr_mul_csr_rowlist_stream): every rank gets its share of hub rows and of tail rows, and no launch tail is paid
    per block.  (Rows r, r+N, ... would not do: the even rows of an un-permuted R-MAT hold 3/4 of the entries.)  Rank 0 then times the whole problem alone: the 1-GPU figure of the
    same box and run."""
    out = {}
    try:
        wname = workload_name("rmat", scale=scale)
        eng.trim_pool()
        dA, t_bcast = shared_operand("rmat", scale=scale)
        rows = dA.dev.row
        my_rows = eng.row_share(dA, dA, world, rank)          # rows by decreasing products, dealt in snake order
        n_mine = int(my_rows.numel())

        def step():
            return eng.csr_mul_csr_rowlist_stream(dA, dA, my_rows.data_ptr(), n_mine)

        ms_step, stats, launches, n_warm = B.timed(step, 2 if one_gpu else 1, 1, max_warm=3 if one_gpu else 1)
        st = stats[-1]
        products, nnz_c = B.all_sum(st["products"], st["nnz"])
        busy = torch.zeros(world, dtype=torch.float64, device="cuda")
        busy[rank] = st["ms_total"]                      # this rank's device time of the last step
        dist.all_reduce(busy, op=dist.ReduceOp.SUM)
        out = {"workload": wname, "scaling": "strong", "n_gpus": world, "ms_per_step": ms_step, "value": 2.0 * products / (ms_step * 1e6),
               "unit": "GFLOP/s", "products": products, "nnz_C": nnz_c, "per_rank_busy_ms": [round(float(x), 1) for x in busy.tolist()],
               "broadcast_ms": t_bcast, "partition": "rows sorted by products, dealt in snake order (ias_row_share + ias_csr_mul_csr_rowlist_stream)",
               "streaming_batches_rank0": st.get("batches"),
               "e2e": {"ms_per_step": ms_step + t_bcast, "value": 2.0 * products / ((ms_step + t_bcast) * 1e6), "unit": "GFLOP/s",
                       "note": "operand resident on rank 0 -> NCCL broadcast of B (3 arrays) + multiply; C is reduced on device "
                               "(nnz, checksum, structure hash), a few scalars come back"}}
        dist.barrier()
        if rank == 0 and not one_gpu:
            # the whole problem on one GPU takes minutes at this scale: the serial equivalent is the sum of the ranks'
            # device times (what one GPU would spend on the N shares one after the other)
            b = out["per_rank_busy_ms"]
            out["one_gpu_same_run"] = None
            out["serial_equivalent_ms"] = round(float(np.sum(b)), 1)
            out["speedup_vs_serial_equivalent"] = float(np.sum(b)) / ms_step
            out["efficiency"] = out["speedup_vs_serial_equivalent"] / world
            out["limiter"] = "slowest rank %.0f ms vs mean %.0f ms of device time; step %.0f ms" % (max(b), float(np.mean(b)), ms_step)
            out["steps"], out["warmup"] = 1, n_warm
        if rank == 0 and one_gpu:
            world_saved, B.world = B.world, 1
            try:
                one = B.csr_leg(dA, 0, rows, 1, 1, True, 0, wname)
            finally:
                B.world = world_saved
            out["one_gpu_same_run"] = {"ms_per_step": one["ms_per_step"], "value": one["value"]}
            out["speedup_vs_one_gpu"] = one["ms_per_step"] / ms_step
            out["efficiency"] = out["speedup_vs_one_gpu"] / world
            b = out["per_rank_busy_ms"]
            out["limiter"] = "slowest rank %.0f ms vs mean %.0f ms of device time; step %.0f ms" % (max(b), float(np.mean(b)), ms_step)
        dist.barrier()
        dA.close()
    except Exception as ex:
        out["error"] = ("%s: %s" % (type(ex).__name__, ex))[:300]
    return out


def multi_gpu_e2e(B, eng, torch, dist, dA, fmt, r0, r1, rank, world, products_total, steps):
    """End to end at N GPUs with host operands: rank r holds rows [r0, r1) of the operand in pinned host memory, uploads
    them (1/N of B per rank over PCIe), the ranks all-gather B over NVLink, each multiplies its row block (CSR path)
    and downloads its block of C.  Timed on the device, max over ranks."""
    try:
        rows, cols, nnz = dA.dev.row, dA.dev.col, dA.dev.nnz
        # host copy of this rank's rows (pinned)
        a_rp = np.zeros(2, dtype=np.int32)
        eng.copy(a_rp.ctypes.data, dA.dev.row_ind_dev + 4 * r0, 4, 1)
        eng.copy(a_rp.ctypes.data + 4, dA.dev.row_ind_dev + 4 * r1, 4, 1)
        s, e = int(a_rp[0]), int(a_rp[1])
        h_len = torch.empty(r1 - r0, dtype=torch.int32).pin_memory()
        h_ci = torch.empty(e - s, dtype=torch.int32).pin_memory()
        h_v = torch.empty(e - s, dtype=torch.float64).pin_memory()
        d_rp = torch.empty(r1 - r0 + 1, dtype=torch.int32, device="cuda")
        eng.copy(d_rp.data_ptr(), dA.dev.row_ind_dev + 4 * r0, 4 * (r1 - r0 + 1), 2)
        h_len.copy_((d_rp[1:] - d_rp[:-1]).cpu())
        eng.copy(h_ci.data_ptr(), dA.dev.col_ind_dev + 4 * s, 4 * (e - s), 1)
        eng.copy(h_v.data_ptr(), dA.dev.values_dev + 8 * s, 8 * (e - s), 1)
        # sizes of every rank's share (metadata, exchanged once)
        meta = torch.zeros(world, 2, dtype=torch.int64, device="cuda")
        meta[rank, 0], meta[rank, 1] = r1 - r0, e - s
        dist.all_reduce(meta, op=dist.ReduceOp.SUM)
        nrows_r, nnz_r = [int(x) for x in meta[:, 0].tolist()], [int(x) for x in meta[:, 1].tolist()]
        g_len = torch.empty(rows, dtype=torch.int32, device="cuda")
        g_ci = torch.empty(nnz, dtype=torch.int32, device="cuda")
        g_v = torch.empty(nnz, dtype=torch.float64, device="cuda")
        g_rp = torch.zeros(rows + 1, dtype=torch.int32, device="cuda")
        len_parts = list(torch.split(g_len, nrows_r))
        ci_parts = list(torch.split(g_ci, nnz_r))
        v_parts = list(torch.split(g_v, nnz_r))
        out_rp = out_ci = out_v = None
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d2h_bytes = 0

        def one():
            nonlocal out_rp, out_ci, out_v, d2h_bytes
            len_parts[rank].copy_(h_len, non_blocking=True)             # H2D: this rank's 1/N
            ci_parts[rank].copy_(h_ci, non_blocking=True)
            v_parts[rank].copy_(h_v, non_blocking=True)
            dist.all_gather(len_parts, len_parts[rank])                  # NVLink
            dist.all_gather(ci_parts, ci_parts[rank])
            dist.all_gather(v_parts, v_parts[rank])
            torch.cumsum(g_len, 0, out=g_rp[1:])
            dB = eng.wrap_device(rows, cols, nnz, g_rp.data_ptr(), g_ci.data_ptr(), g_v.data_ptr())
            c64, st = eng.CSR_MUL_CSR_DEV(dB, dB, rows=(r0, r1), keep=True)
            if out_ci is None or out_ci.numel() < c64.nnz:
                out_rp = torch.empty(c64.row + 1, dtype=torch.int64).pin_memory()
                out_ci = torch.empty(c64.nnz, dtype=torch.int32).pin_memory()
                out_v = torch.empty(c64.nnz, dtype=torch.float64).pin_memory()
            eng.copy(out_rp.data_ptr(), c64.row_ptr_dev, 8 * (c64.row + 1), 1)      # D2H: this rank's block of C
            eng.copy(out_ci.data_ptr(), c64.col_ind_dev, 4 * c64.nnz, 1)
            eng.copy(out_v.data_ptr(), c64.values_dev, 8 * c64.nnz, 1)
            d2h_bytes = 8 * (c64.row + 1) + 12 * c64.nnz
            eng.free_csr64(c64)
            dB.close()
            return float(out_v[-1]) if c64.nnz else 0.0

        one()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(steps):
            one()
        ev1.record()
        torch.cuda.synchronize(); dist.barrier()
        ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3) / steps
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tot = torch.tensor([4 * (r1 - r0) + 12 * (e - s), d2h_bytes], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        return {"value": 2.0 * products_total / (ms * 1e6), "unit": "GFLOP/s", "ms_per_step": ms,
                "h2d_bytes_per_step": int(tot[0].item()), "d2h_bytes_per_step": int(tot[1].item()),
                "api": "per rank: pinned host rows -> H2D of 1/N of B -> NCCL all_gather over NVLink -> ias_csr_mul_csr_rows_dev64 -> D2H of its C block"}
    except Exception as ex:
        return {"value": None, "error": ("%s: %s" % (type(ex).__name__, ex))[:300]}


if __name__ == "__main__":
    main()
