#!/usr/bin/env python
"""bench.py -- the measurement contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload banded|cant|rmat|laplace-iter] [--dtype f32|f64] [--rows-per-gpu R]

One "step" = one pass of the hot path over the workload = one SpMV in each of the reference's five
formats (COO, CSR, ELL, SELL-32, CMRS) on the same matrix.  (--workload cant: the cant-shaped matrix
fits in L2, so 7 independent copies are used in rotation and the K steps are one CUDA-graph replay;
--workload rmat / laplace-iter: BASELINE configs[3] / configs[4])  Default workload: BASELINE.json
configs[2], the synthetic banded FEM-like matrix, 2 097 152 rows x 64 nnz/row per GPU, fp32,
generated on the device (weak scaling: rank r owns rows [r*R, (r+1)*R) of the (N*R)-row global
matrix, x replicated, no collective in the data path).  Every format's arrays (0.8-1.6 GB) are far
larger than the 126 MB L2, so no flush is needed between iterations.

Prints ONE JSON line.  metric = aggregate SpMV GFLOP/s (2*nnz flops per SpMV, the reference's own
FLOP model, inc/helper_functions.h:171-172) over the five formats; `formats` carries the per-format
GFLOP/s, algorithmic GB/s and fraction of the measured HBM peak; `roofline` describes the dominant
(slowest) kernel; `cpu_baseline` is the oracle port on the host cores on a bounded sample.

--impl reference times the reference's own CPU code (oracle/_ref, compute_using_cpu built -O3 from
the unmodified sources; fp64 because the reference has no fp32) on the box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

os.environ.setdefault("OMP_WAIT_POLICY", "passive")
os.environ.setdefault("OMP_PROC_BIND", "close")

import numpy as np  # noqa: E402

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FORMATS = ("coo", "csr", "ell", "sell", "cmrs")

# stdout must carry exactly ONE JSON line, but libraries write there too (NCCL prints its version
# line, the reference's compute_using_cpu printf()s its timing block): everything written to fd 1
# during the run goes to stderr, and only emit_json() writes to the real stdout.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_json(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    fd = _REAL_STDOUT if _REAL_STDOUT is not None else 1
    while line:
        line = line[os.write(fd, line):]
NOMINAL_HBM_GBS = 8000.0   # BASELINE.json north_star
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region (pynvml, 10 ms period)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "app_clocks", 0x100: "display",
               0x10: "sync_boost"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.ok:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join(timeout=1)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------
BANDED = dict(nnz_per_row=64, half_band=2000, seed=42, x_seed=7)


def build_banded_device(pkg, ctx, n_global, row_begin, row_count, dtype):
    L = pkg.lib()
    nnz = L.b200_gen_banded_nnz(n_global, row_begin, row_count, BANDED["nnz_per_row"])
    rows, cols, vals = ctx.empty(nnz, np.int32), ctx.empty(nnz, np.int32), ctx.empty(nnz, np.float64)
    pkg.check(L.b200_gen_banded_coo(ctx.h, n_global, row_begin, row_count, BANDED["nnz_per_row"],
                                    BANDED["half_band"], BANDED["seed"], rows.ptr, cols.ptr, vals.ptr),
              "b200_gen_banded_coo")
    # shard-local row indices: each rank builds its formats on its own row block; columns stay global
    pkg.check(L.b200_offset_i32(ctx.h, rows.ptr, nnz, -row_begin), "b200_offset_i32")
    coo = pkg.CooMatrix(ctx, row_count, n_global, rows, cols, vals)
    x = ctx.empty(n_global, dtype)
    if np.dtype(dtype) == np.float32:
        pkg.check(L.b200_gen_uniform_f32(ctx.h, x.ptr, n_global, BANDED["x_seed"], 0.0, 1.0), "gen x")
    else:
        pkg.check(L.b200_gen_uniform_f64(ctx.h, x.ptr, n_global, BANDED["x_seed"], 0.0, 1.0), "gen x")
    return coo, x


def banded_host(pkg, n_global, row_begin, row_count):
    L = pkg.lib()
    nnz = row_count * BANDED["nnz_per_row"]
    rows, cols, vals = np.empty(nnz, np.int32), np.empty(nnz, np.int32), np.empty(nnz, np.float64)
    pkg.check(L.b200_gen_banded_coo_host(n_global, row_begin, row_count, BANDED["nnz_per_row"],
                                         BANDED["half_band"], BANDED["seed"], rows.ctypes.data,
                                         cols.ctypes.data, vals.ctypes.data), "gen host")
    x = np.empty(n_global, np.float64)
    pkg.check(L.b200_gen_uniform_f64_host(x.ctypes.data, n_global, BANDED["x_seed"], 0.0, 1.0), "gen x host")
    return rows, cols, vals, x


def cant_host():
    """cant-shaped stand-in parsed from generated MatrixMarket text (row-sorted file)."""
    import subprocess
    import tempfile
    import pandas as pd
    gen = ROOT / "opencl-spmv-algorithms_b200" / "tools" / "gen_mtx"
    with tempfile.TemporaryDirectory() as d:
        path = Path(d) / "cant-sorted.mtx"
        subprocess.run([str(gen), "--order", "row", "--out", str(path)], check=True)
        with open(path) as f:  # banner, one comment line, size line (tools/gen_mtx.c)
            f.readline()
            f.readline()
            n_rows, n_cols, _ = (int(t) for t in f.readline().split())
        t = pd.read_csv(path, sep=" ", skiprows=3, header=None, names=["r", "c", "v"],
                        dtype={"r": np.int32, "c": np.int32, "v": np.float64})
    rows, cols, vals = (t.r.values - 1).astype(np.int32), (t.c.values - 1).astype(np.int32), t.v.values
    return n_rows, n_cols, rows, cols, vals, np.arange(n_cols, dtype=np.float64)


# --------------------------------------------------------------------------------------------
# CPU arms (oracle port / compiled reference).  The ONLY place bench.py touches oracle/.
# --------------------------------------------------------------------------------------------
class quiet_stdout:
    """The reference's compute_using_cpu printf()s its own timing block; keep our stdout to the
    single JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        import ctypes
        ctypes.CDLL(None).fflush(None)
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_arm(n_rows, n_cols, rows, cols, vals, x, dtype, kind, reps, warm):
    """Time one SpMV per format on the host cores; returns (seconds per five-format pass, details).
    kind='port': oracle/liboracle.so (-O3, OpenMP) in `dtype`.
    kind='reference': the reference's own compute_using_cpu (fp64, -O3 build of the unmodified
    sources) for coo/csr/ell/cmrs; sigma_c.c has no CPU path, so SELL uses the port."""
    from oracle import binding as O
    threads = os.cpu_count() or 1
    O.lib().orc_set_threads(threads)
    nnz = len(rows)
    ptr, _ = O.build_csr(n_rows, rows)
    K = int(np.bincount(rows, minlength=n_rows).max())
    ec, ed = O.build_ell(n_rows, K, rows, cols, vals)
    ri, sc, sd = O.build_sell(n_rows, rows, cols, vals)
    sp, ris = O.build_cmrs(n_rows, rows)
    dt = np.float64 if kind == "reference" else dtype
    v, xx = vals.astype(dt), x.astype(dt)
    ed_t, sd_t = ed.astype(dt), sd.astype(dt)
    use_ref = kind == "reference" and O.ref_available()
    if use_ref:
        arrs = {
            "coo": dict(rows=rows, cols=cols, data=v, vect=xx),
            "csr": dict(ptr=ptr, cols=cols, data=v, vect=xx),
            "ell": dict(cols=ec, data=ed_t, vect=xx),
            "cmrs": dict(strip_ptr=sp, row_in_strip=ris, cols=cols, data=v, vect=xx),
        }
        runs = {
            "coo": lambda: O.ref_compute_using_cpu("coo", arrs["coo"], n_rows, nnz, opt=True),
            "csr": lambda: O.ref_compute_using_cpu("csr", arrs["csr"], n_rows, nnz, opt=True),
            "ell": lambda: O.ref_compute_using_cpu("ell", arrs["ell"], n_rows, nnz, opt=True, row_size=K),
            "cmrs": lambda: O.ref_compute_using_cpu("cmrs", arrs["cmrs"], n_rows, nnz, opt=True),
            "sell": lambda: O.spmv_sell(ri, sc, sd_t, xx, dtype=dt),
        }
    else:
        runs = {
            "coo": lambda: O.spmv_coo(n_rows, rows, cols, v, xx, dtype=dt),
            "csr": lambda: O.spmv_csr(n_rows, ptr, cols, v, xx, dtype=dt),
            "ell": lambda: O.spmv_ell(n_rows, K, ec, ed_t, xx, dtype=dt),
            "sell": lambda: O.spmv_sell(ri, sc, sd_t, xx, dtype=dt),
            "cmrs": lambda: O.spmv_cmrs(n_rows, sp, ris, cols, v, xx, dtype=dt),
        }
    per = {}
    with quiet_stdout():
        for f in FORMATS:
            for _ in range(warm):
                runs[f]()
            best = float("inf")
            for _ in range(reps):
                t0 = time.perf_counter()
                runs[f]()
                best = min(best, time.perf_counter() - t0)
            per[f] = best
        # the same symbol built with the reference's OWN flags (no -O,
        # Makefile:18) -- its first call (what the "CPU calculations" block of a run reports, minus the
        # OpenMP team start-up that the -O3 runs above have already paid), then warm
        as_shipped = None
        if use_ref:
            as_shipped = {}
            unopt = {
                "coo": lambda: O.ref_compute_using_cpu("coo", arrs["coo"], n_rows, nnz, opt=False),
                "csr": lambda: O.ref_compute_using_cpu("csr", arrs["csr"], n_rows, nnz, opt=False),
                "ell": lambda: O.ref_compute_using_cpu("ell", arrs["ell"], n_rows, nnz, opt=False, row_size=K),
                "cmrs": lambda: O.ref_compute_using_cpu("cmrs", arrs["cmrs"], n_rows, nnz, opt=False),
            }
            for f, run in unopt.items():
                t0 = time.perf_counter()
                run()
                cold = time.perf_counter() - t0
                best = float("inf")
                for _ in range(max(reps, 2)):
                    t0 = time.perf_counter()
                    run()
                    best = min(best, time.perf_counter() - t0)
                as_shipped[f] = {"first_call_ms": round(cold * 1e3, 3), "warm_best_ms": round(best * 1e3, 3),
                                 "gflops_first_call": round(2.0 * nnz / cold * 1e-9, 3),
                                 "gflops_warm": round(2.0 * nnz / best * 1e-9, 3)}
    total = sum(per.values())
    detail = {f: round(2.0 * nnz / per[f] * 1e-9, 3) for f in FORMATS}
    if as_shipped is not None:
        detail = dict(detail)
        cpu_arm.as_shipped = as_shipped
    return total, detail, threads, ("reference" if use_ref else "port"), ("f32" if np.dtype(dt) == np.float32 else "f64")


# --------------------------------------------------------------------------------------------
def main():
    capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="banded", choices=["banded", "cant", "laplace-iter", "rmat"])
    ap.add_argument("--rmat-scale", type=int, default=24)
    ap.add_argument("--rmat-edge-factor", type=int, default=16)
    ap.add_argument("--rmat-max-sell-bytes", type=float, default=24e9,
                    help="rmat: skip running a SELL variant whose padded arrays exceed this")
    ap.add_argument("--grid", type=int, default=400, help="laplace-iter: nx = ny")
    ap.add_argument("--nz-per-gpu", type=int, default=50, help="laplace-iter: z planes per rank")
    ap.add_argument("--iter-format", default="csr", choices=["csr", "sell"])
    ap.add_argument("--dtype", default=None, choices=["f32", "f64"])
    ap.add_argument("--rows-per-gpu", type=int, default=2097152)
    ap.add_argument("--cant-copies", type=int, default=7,
                    help="cant: independent copies of the format arrays used in rotation (cold L2 without a flush)")
    ap.add_argument("--cpu-sample-rows", type=int, default=262144)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dtype = np.dtype(np.float32 if (args.dtype or ("f32" if args.workload == "banded" else "f64")) == "f32"
                     else np.float64)
    dname = "f32" if dtype == np.float32 else "f64"

    from __graft_entry__ import load_package
    pkg = load_package()

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(pkg, args, dtype)
    if args.workload == "laplace-iter":
        return laplace_iter_arm(pkg, args, rank, world, local_rank)
    if args.workload == "rmat":
        return rmat_arm(pkg, args, rank, world, local_rank)

    # ---------------- the B200 arm ----------------
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = pkg.Context(local_rank)
    L = pkg.lib()

    if args.workload == "banded":
        R = args.rows_per_gpu
        n_global = R * world
        row_begin = rank * R
        coo, x = build_banded_device(pkg, ctx, n_global, row_begin, R, dtype)
        n_rows, n_cols = R, n_global
        workload = (f"banded FEM-like (BASELINE configs[2]): {R} rows x {BANDED['nnz_per_row']} nnz/row per GPU, "
                    f"global {n_global} x {n_global}, half-band {BANDED['half_band']}, device-generated")
    else:
        n_rows, n_cols, rows_h, cols_h, vals_h, x_h = cant_host()
        coo = pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows_h, cols_h, vals_h)
        x = ctx.array(x_h.astype(dtype))
        workload = "cant-shaped stand-in (BASELINE configs[1]): 62451 x 62451, 4325625 nnz, x = ramp"
    nnz = coo.nnz

    def build_set(coo_m):
        allm = pkg.build_all(coo_m, dtype)
        allm["csr"].plan()
        # "ell" = the kernel on the reference's row-major arrays (fastest ELL kernel on B200); the
        # column-major thread-per-row kernel is measured next to it
        # "cmrs_packed" = the derived 4+V bytes/entry layout (row_in_strip folded into the column word),
        # measured next to the reference two-array layout, never in the headline
        return ({"coo": allm["coo"], "csr": allm["csr"], "ell": allm["ell"], "sell": allm["sell"],
                 "cmrs": allm["cmrs"]}, {"ell_colmajor": allm["ellcm"], "cmrs_packed": allm["cmrs"].packed()})

    # The cant-shaped formats (53-70 MB each) fit in the 126 MB L2.  "Inputs larger than L2" is
    # restored by ROTATION: n_copies independent copies of every format's arrays (own COO triples,
    # nothing shared), launch i reading copy i mod n_copies -- arrays last touched n_copies launches
    # (> 370 MB of other traffic) earlier.  No flush kernel runs inside the timed region, so the L2
    # holds clean lines and the kernels run back to back exactly as in the banded workload.
    n_copies = max(1, args.cant_copies) if args.workload == "cant" else 1
    sets = [build_set(coo)]
    for _ in range(n_copies - 1):
        sets.append(build_set(pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows_h, cols_h, vals_h)))
    mats, extra = sets[0]
    y = {f: ctx.zeros(n_rows, dtype) for f in list(mats) + list(extra)}
    bytes_alg = {f: m.nbytes(dtype) for f, m in {**mats, **extra}.items()}
    ctx.set_l2_persist(x)
    ctx.sync()
    launch_no = [0]

    def next_set():
        launch_no[0] += 1
        return sets[launch_no[0] % n_copies]

    # old methodology, kept as context for the L2-sized workload only: a 256 MiB memset before every
    # launch (leaves the L2 full of DIRTY lines whose write-back competes with the kernel's reads)
    flush = ctx.empty(256 << 20, np.uint8) if args.workload == "cant" else None

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

    overlap = None
    if n_copies == 1:
        for _ in range(args.warmup):
            for f in mats:
                mats[f].spmv(x, y[f])
        barrier()

        # timed region: K steps; per-format events inside, whole-region events outside
        ev = {f: [(ctx.event(), ctx.event()) for _ in range(args.steps)] for f in mats}
        e0, e1 = ctx.event(), ctx.event()
        with ClockSampler(local_rank) as clk:
            barrier()
            e0.record()
            for s in range(args.steps):
                for f, m in mats.items():
                    ev[f][s][0].record()
                    m.spmv(x, y[f])
                    ev[f][s][1].record()
            e1.record()
            barrier()
        total_ms = e0.elapsed_ms_until(e1)
        per_ms = {f: float(np.mean([a.elapsed_ms_until(b) for a, b in ev[f]])) for f in mats}
        step_ms = total_ms / args.steps
    else:
        # A cant-sized SpMV lasts ~10 us, less than this Python loop needs to issue one launch, so the
        # launches are recorded into CUDA graphs (b200_graph_*) and replayed:
        #   g_steps = exactly K steps (5K launches, rotating over the copies): the timed region is ONE
        #             replay of it -> ms_per_step, value;
        #   g_fmt[f] = 4*n_copies launches of format f alone, rotating -> per-format time.
        def record_steps(k):
            with ctx.record_graph() as g:
                for _ in range(k):
                    for f in mats:
                        next_set()[0][f].spmv(x, y[f])
            return g

        def record_format(f, which, n):
            with ctx.record_graph() as g:
                for _ in range(n):
                    next_set()[which][f].spmv(x, y[f])
            return g

        for f in mats:  # un-graphed pass over every copy first: loads the kernels, warms the plans
            for st in sets:
                st[0][f].spmv(x, y[f])
        g_steps = record_steps(args.steps)
        g_warm = record_steps(max(args.warmup, n_copies))
        n_f = 4 * n_copies
        g_fmt = {f: record_format(f, 0, n_f) for f in mats}
        e0, e1 = ctx.event(), ctx.event()
        with ClockSampler(local_rank) as clk:
            g_warm.launch()
            barrier()
            e0.record()
            g_steps.launch()
            e1.record()
            barrier()
            total_ms = e0.elapsed_ms_until(e1)
            per_ms = {}
            for f in mats:
                g_fmt[f].launch()
                reps = []
                for _ in range(5):
                    a, b = ctx.event(), ctx.event()
                    a.record()
                    g_fmt[f].launch()
                    b.record()
                    ctx.sync()
                    reps.append(a.elapsed_ms_until(b) / n_f)
                per_ms[f] = float(np.mean(reps))
        step_ms = total_ms / args.steps

        # the same graphs recorded with launch overlap on (b200_ctx_set_launch_overlap: programmatic
        # dependent launch -- every kernel streams its matrix arrays while its predecessor drains and
        # touches x / y only after it has finished).  Reported next to the in-order numbers.
        ctx.set_launch_overlap(True)
        g_steps_o = record_steps(args.steps)
        g_fmt_o = {f: record_format(f, 0, n_f) for f in mats}
        ctx.set_launch_overlap(False)
        g_steps_o.launch()
        barrier()
        e0.record()
        g_steps_o.launch()
        e1.record()
        barrier()
        overlap = {"ms_per_step": e0.elapsed_ms_until(e1) / args.steps, "formats": {}}
        for f in mats:
            g_fmt_o[f].launch()
            reps = []
            for _ in range(5):
                a, b = ctx.event(), ctx.event()
                a.record()
                g_fmt_o[f].launch()
                b.record()
                ctx.sync()
                reps.append(a.elapsed_ms_until(b) / n_f)
            overlap["formats"][f] = float(np.mean(reps))

    # extra (untimed for the headline): the column-major ELL kernel, same method
    for f in extra:
        if n_copies == 1:
            for _ in range(3):
                extra[f].spmv(x, y[f])
            pairs = [(ctx.event(), ctx.event()) for _ in range(10)]
            for a, b in pairs:
                a.record()
                extra[f].spmv(x, y[f])
                b.record()
            ctx.sync()
            per_ms[f] = float(np.mean([a.elapsed_ms_until(b) for a, b in pairs]))
        else:
            for st in sets:
                st[1][f].spmv(x, y[f])
            g = record_format(f, 1, n_f)
            g.launch()
            a, b = ctx.event(), ctx.event()
            a.record()
            g.launch()
            b.record()
            ctx.sync()
            per_ms[f] = a.elapsed_ms_until(b) / n_f

    # context for the L2-sized workload, never in the headline: (1) the previous methodology, a
    # 256 MiB memset before every launch; (2) the same arrays back to back, L2-resident (what an
    # iterative solver sees after its first step)
    warm = flushed = None
    if flush is not None:
        warm, flushed = {}, {}
        for f, m in {**mats, **extra}.items():
            pairs = [(ctx.event(), ctx.event()) for _ in range(20)]
            for a, b in pairs:
                flush.fill_bytes(1)
                a.record()
                m.spmv(x, y[f])
                b.record()
            ctx.sync()
            ms = float(np.mean([a.elapsed_ms_until(b) for a, b in pairs]))
            flushed[f] = {"ms": round(ms, 5), "gbs_algorithmic": round(bytes_alg[f] / (ms * 1e-3) * 1e-9, 1)}
            for _ in range(3):
                m.spmv(x, y[f])
            a, b = ctx.event(), ctx.event()
            a.record()
            for _ in range(20):
                m.spmv(x, y[f])
            b.record()
            ctx.sync()
            ms = a.elapsed_ms_until(b) / 20
            warm[f] = {"ms": round(ms, 5), "gbs_algorithmic": round(bytes_alg[f] / (ms * 1e-3) * 1e-9, 1)}

    if dist is not None:
        import torch
        t = torch.tensor([step_ms] + [per_ms[f] for f in mats], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms = float(t[0])
        for i, f in enumerate(mats):
            per_ms[f] = float(t[1 + i])

    peak, peak_src = measured_peak()
    flops_step = 2.0 * nnz * len(mats) * world
    value = flops_step / (step_ms * 1e-3) * 1e-9
    fm = {}
    for f in list(mats) + list(extra):
        gbs = bytes_alg[f] / (per_ms[f] * 1e-3) * 1e-9
        fm[f] = {"ms": round(per_ms[f], 5), "gflops": round(2.0 * nnz / (per_ms[f] * 1e-3) * 1e-9, 2),
                 "alg_bytes": int(bytes_alg[f]), "gbs": round(gbs, 1),
                 "frac_measured": round(gbs / peak, 4), "frac_nominal_8TBs": round(gbs / NOMINAL_HBM_GBS, 4)}
    dom = max(mats, key=lambda f: per_ms[f])
    # DRAM bytes per launch of that kernel from the committed ncu --set full capture of this very
    # workload (profiles/ncu_traffic.json); null for workloads that were not captured
    traffic = None
    try:
        cap = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
        for c in cap if isinstance(cap, list) else [cap]:
            if c["workload"] == args.workload and c["dtype"] == dname and world == 1 and c.get("rows", 2097152) == n_rows:
                traffic = c["traffic_bytes"].get(dom)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": fm[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": fm[dom]["frac_measured"], "traffic": traffic, "peak_source": peak_src,
                "per_format_frac": {f: fm[f]["frac_measured"] for f in mats}}

    # ---------------- e2e: host x in, host y out, through the C ABI, matrix resident ----------
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        V = dtype.itemsize
        hx, hy = C.c_void_p(), C.c_void_p()
        pkg.check(L.b200_host_alloc_pinned(n_cols * V, C.byref(hx)), "pinned x")
        pkg.check(L.b200_host_alloc_pinned(n_rows * V, C.byref(hy)), "pinned y")
        x_host = np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_byte)), shape=(n_cols * V,)).view(dtype)
        x_host[:] = x.download()
        xin = ctx.empty(n_cols, dtype)
        steps_e = max(3, min(args.steps, 20))

        def e2e_step():
            for f, m in mats.items():
                pkg.check(L.b200_memcpy_h2d_async(ctx.h, xin.ptr, hx, n_cols * V), "h2d x")
                m.spmv(xin, y[f])
                pkg.check(L.b200_memcpy_d2h_async(ctx.h, hy, y[f].ptr, n_rows * V), "d2h y")
            ctx.sync()

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        a, b = ctx.event(), ctx.event()
        a.record()
        for _ in range(steps_e):
            e2e_step()
        b.record()
        barrier()
        e2e_ms = a.elapsed_ms_until(b) / steps_e
        wall_ms = (time.perf_counter() - t0) * 1e3 / steps_e
        if dist is not None:
            import torch
            t = torch.tensor([e2e_ms, wall_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms, wall_ms = float(t[0]), float(t[1])
        # the same five calls issued on TWO in-order queues (contexts) of the same device, formats
        # alternating: PCIe is full duplex, so the y download of one call overlaps the x upload and
        # kernel of the next.  Wall clock with both queues drained on both sides (events live on one
        # stream only).
        ctx2 = pkg.Context(local_rank)
        hx2, hy2 = C.c_void_p(), C.c_void_p()
        pkg.check(L.b200_host_alloc_pinned(n_cols * V, C.byref(hx2)), "pinned x2")
        pkg.check(L.b200_host_alloc_pinned(n_rows * V, C.byref(hy2)), "pinned y2")
        C.memmove(hx2, hx, n_cols * V)
        xin2 = ctx2.empty(n_cols, dtype)
        queues = [(ctx, xin, hx, hy), (ctx2, xin2, hx2, hy2)]
        owner = {f: queues[i % 2] for i, f in enumerate(mats)}

        def e2e_step2():
            for f, m in mats.items():
                q, xq, hxq, hyq = owner[f]
                m.ctx = q
                pkg.check(L.b200_memcpy_h2d_async(q.h, xq.ptr, hxq, n_cols * V), "h2d x")
                m.spmv(xq, y[f])
                pkg.check(L.b200_memcpy_d2h_async(q.h, hyq, y[f].ptr, n_rows * V), "d2h y")
            ctx.sync()
            ctx2.sync()

        for _ in range(2):
            e2e_step2()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps_e):
            e2e_step2()
        barrier()
        wall2_ms = (time.perf_counter() - t0) * 1e3 / steps_e
        for m in mats.values():
            m.ctx = ctx
        if dist is not None:
            import torch
            t = torch.tensor([wall2_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall2_ms = float(t[0])
        best_ms = min(e2e_ms, wall2_ms)
        e2e = {"value": round(flops_step / (best_ms * 1e-3) * 1e-9, 2), "unit": "GFLOP/s",
               "h2d_bytes_per_step": int(len(mats) * n_cols * V * world),
               "d2h_bytes_per_step": int(len(mats) * n_rows * V * world),
               "ms_per_step": round(best_ms, 4), "steps": steps_e,
               "one_queue": {"ms_per_step": round(e2e_ms, 4), "wall_ms_per_step": round(wall_ms, 4),
                             "gflops": round(flops_step / (e2e_ms * 1e-3) * 1e-9, 2)},
               "two_queues": {"wall_ms_per_step": round(wall2_ms, 4),
                              "gflops": round(flops_step / (wall2_ms * 1e-3) * 1e-9, 2)},
               "what": "per format: pinned-host x -> device, SpMV through the C ABI, y -> pinned host; format "
                       "arrays uploaded once before the timed region, as the reference driver does (csr.c:183-193). "
                       "value = the better of one in-order queue (CUDA events) and two queues with alternating "
                       "formats (wall clock, both drained)"}
        ctx2.sync()
        del xin2
        L.b200_host_free_pinned(hx)
        L.b200_host_free_pinned(hy)
        L.b200_host_free_pinned(hx2)
        L.b200_host_free_pinned(hy2)
        ctx2.close()

    # ---------------- CPU baseline (rank 0, N=1 only): oracle port on a bounded sample ---------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if args.workload == "banded":
            sr = min(args.cpu_sample_rows, n_rows)
            rows_h, cols_h, vals_h, x_h = banded_host(pkg, n_cols, 0, sr)
            sample = f"first {sr} rows ({sr * BANDED['nnz_per_row']} nnz) of the same banded matrix, five formats, best of 5"
        else:
            sr = n_rows
            sample = "the whole cant-shaped matrix, five formats, best of 5"
        sec, detail, threads, kind, cdt = cpu_arm(sr, n_cols, rows_h, cols_h, vals_h, x_h, dtype, "port", 5, 2)
        cpu = {"value": round(2.0 * len(rows_h) * len(FORMATS) / sec * 1e-9, 3), "unit": "GFLOP/s",
               "cores": threads, "cpu_model": cpu_model(), "omp_wait_policy": os.environ.get("OMP_WAIT_POLICY"),
               "kind": kind, "dtype": cdt, "sample": sample, "per_format_gflops": detail}

    if rank == 0:
        out = {
            "metric": "SpMV GFLOP/s, aggregate over the five formats (COO, CSR, ELL, SELL-32, CMRS); "
                      "per-format GFLOP/s and HBM GB/s (% of peak) in `formats`",
            "value": round(value, 2), "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(step_ms, 5), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dname, "data": "synthetic",
            "config": {"workload": workload, "formats": list(mats), "nnz_per_gpu": int(nnz),
                       "rows_per_gpu": int(n_rows), "cols": int(n_cols),
                       "partition": f"row blocks, {world} rank(s), x replicated, no collective",
                       "cache": (f"inputs larger than L2 by rotation: {n_copies} independent copies of every format's "
                                 f"arrays ({min(bytes_alg.values()) >> 20}-{max(bytes_alg.values()) >> 20} MiB each), launch i "
                                 f"reads copy i mod {n_copies}; no flush; the K steps are one CUDA-graph replay" if n_copies > 1 else
                                 "inputs larger than L2 (0.8-1.6 GB per format), no flush")},
            "formats": fm,
            "launch_overlap": (None if overlap is None else {
                "what": "same K-step graph with b200_ctx_set_launch_overlap(1): kernels are programmatic dependents "
                        "(matrix arrays streamed while the previous launch drains; x read / y written after it ends)",
                "ms_per_step": round(overlap["ms_per_step"], 5),
                "value": round(flops_step / (overlap["ms_per_step"] * 1e-3) * 1e-9, 2),
                "formats": {f: {"ms": round(ms, 5), "frac_measured": round(bytes_alg[f] / (ms * 1e-3) * 1e-9 / peak, 4)}
                            for f, ms in overlap["formats"].items()}}),
            "formats_flush_each_launch_context_only": flushed,
            "formats_warm_l2_context_only": warm, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(args.steps * len(mats)), "clocks": clk.summary(),
        }
        emit_json(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


def rmat_arm(pkg, args, rank, world, local_rank):
    """BASELINE configs[3]: power-law R-MAT (scale 24: 16.8 M rows, ~270 M nnz after duplicate
    removal), fp32, CSR vs SELL-32-sigma sweep (+ COO, CMRS), rows partitioned nnz-balanced over the
    ranks (STRONG scaling: the global matrix is fixed), x replicated, no collective."""
    import ctypes as C
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = pkg.Context(local_rank)
    L = pkg.lib()
    scale, ef, abc, seed = args.rmat_scale, args.rmat_edge_factor, (0.57, 0.19, 0.19), 5
    n = 1 << scale
    dtype = np.dtype(np.float32 if (args.dtype or "f32") == "f32" else np.float64)

    def candidates(r0, cnt):
        c = C.c_longlong(0)
        pkg.check(L.b200_gen_rmat_count(ctx.h, scale, ef, *abc, seed, r0, cnt, C.byref(c)), "rmat count")
        return c.value

    # nnz-balanced cuts (multiples of 32) by bisection on the candidate count of rows [0, r)
    cuts = [0]
    if world > 1:
        total = candidates(0, n)
        for k in range(1, world):
            lo, hi = 0, n // 32
            while lo < hi:
                mid = (lo + hi) // 2
                if candidates(0, mid * 32) < total * k // world:
                    lo = mid + 1
                else:
                    hi = mid
            cuts.append(max(lo * 32, cuts[-1]))
    cuts.append(n)
    r0, r1 = cuts[rank], cuts[rank + 1]
    cap = candidates(r0, r1 - r0)
    rows, cols, vals = ctx.empty(cap, np.int32), ctx.empty(cap, np.int32), ctx.empty(cap, np.float64)
    nnz_c = C.c_longlong(0)
    pkg.check(L.b200_gen_rmat_coo(ctx.h, scale, ef, *abc, seed, r0, r1 - r0, cap, rows.ptr, cols.ptr, vals.ptr,
                                  C.byref(nnz_c)), "rmat gen")
    nnz = nnz_c.value
    pkg.check(L.b200_offset_i32(ctx.h, rows.ptr, nnz, -r0), "rebase rows")
    rows.n = cols.n = vals.n = nnz  # views of the first nnz entries
    n_rows = r1 - r0
    coo = pkg.CooMatrix(ctx, n_rows, n, rows, cols, vals)
    x = ctx.empty(n, dtype)
    gen_x = L.b200_gen_uniform_f32 if dtype == np.float32 else L.b200_gen_uniform_f64
    pkg.check(gen_x(ctx.h, x.ptr, n, 7, 0.0, 1.0), "gen x")
    csr = pkg.CsrMatrix(coo, check_sorted=False)
    info, st = csr.plan_info(), csr.row_stats()
    y = ctx.zeros(n_rows, dtype)

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

    def time_mat(m):
        for _ in range(args.warmup):
            m.spmv(x, y)
        barrier()
        a, b = ctx.event(), ctx.event()
        a.record()
        for _ in range(args.steps):
            m.spmv(x, y)
        b.record()
        barrier()
        return a.elapsed_ms_until(b) / args.steps

    peak, peak_src = measured_peak()
    results, order = {}, []

    def record(name, m, extra=None):
        ms = time_mat(m)
        alg = m.nbytes(dtype)
        results[name] = dict(ms=ms, alg_bytes=int(alg), nnz=int(nnz), **(extra or {}))
        order.append(name)

    with ClockSampler(local_rank) as clk:
        record("csr", csr)
        record("coo", coo)
        record("cmrs", pkg.CmrsMatrix(csr))
        for sigma in (1, 32, 256, 4096, 65536, n_rows):
            total = C.c_longlong(0)
            sp = ctx.empty(L.b200_sell_num_slices(n_rows, 32) + 1, np.int64)
            perm = ctx.empty(n_rows, np.int32)
            pkg.check(L.b200_build_sell_ptr(ctx.h, csr.ptr.ptr, n_rows, 32, sigma, perm.ptr, sp.ptr,
                                            C.byref(total)), "sell ptr")
            name = f"sell_sigma{sigma if sigma != n_rows else 'R'}"
            pad = total.value / max(nnz, 1)
            del sp, perm
            if total.value * (4 + dtype.itemsize) > args.rmat_max_sell_bytes:
                results[name] = dict(ms=None, padding_factor=round(pad, 3), skipped="padded arrays exceed --rmat-max-sell-bytes")
                order.append(name)
                continue
            m = pkg.SellMatrix(csr, dtype, sigma=sigma, wide=True)
            record(name, m, dict(padding_factor=round(pad, 3)))
            del m
    # max over ranks per format, sum of nnz (strong scaling: the job is the whole matrix)
    names = [k for k in order if results[k].get("ms") is not None]
    ms = np.array([results[k]["ms"] for k in names])
    tot = np.array([float(nnz)])
    if dist is not None:
        import torch
        t = torch.tensor(ms, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.cpu().numpy()
        t2 = torch.tensor(tot, device="cuda", dtype=torch.float64)
        dist.all_reduce(t2, op=dist.ReduceOp.SUM)
        tot = t2.cpu().numpy()
    fm = {}
    for k in order:
        r = results[k]
        if r.get("ms") is None:
            fm[k] = r
            continue
        t_ms = float(ms[names.index(k)])
        gbs = r["alg_bytes"] / (r["ms"] * 1e-3) * 1e-9  # this rank's own bytes / own time
        fm[k] = {"ms": round(t_ms, 5), "gflops": round(2.0 * tot[0] / (t_ms * 1e-3) * 1e-9, 2),
                 "alg_bytes_rank0": r["alg_bytes"], "gbs_rank0": round(gbs, 1), "frac_measured_rank0": round(gbs / peak, 4)}
        if "padding_factor" in r:
            fm[k]["padding_factor"] = r["padding_factor"]
    if rank == 0:
        best_sell = min((k for k in names if k.startswith("sell")), key=lambda k: fm[k]["ms"], default=None)
        out = {
            "metric": "SpMV GFLOP/s of CSR on the power-law matrix (2*nnz flops); COO, CMRS and the SELL-32-sigma "
                      "sweep with padding factors in `formats`",
            "value": fm["csr"]["gflops"], "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": fm["csr"]["ms"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if dtype == np.float32 else "f64", "data": "synthetic",
            "config": {"workload": f"R-MAT power-law (BASELINE configs[3]): scale {scale} = {n} rows, edge factor {ef}, "
                                   f"(a,b,c,d)=(0.57,0.19,0.19,0.05), diagonal added, duplicates removed: {int(tot[0])} nnz; "
                                   f"{world} nnz-balanced row block(s), x replicated, no collective",
                       "rank0": {"rows": int(n_rows), "nnz": int(nnz), "mean_len": round(info.mean_len, 2),
                                 "max_len": int(st.max_len), "csr_kernel": "nnz-split stream" if info.stream_tiles else
                                 f"vector, {info.lanes_per_row} lanes/row + {info.n_long_rows} long rows"},
                       "ell": f"not run: K = longest row = {int(st.max_len)} would need {n_rows * st.max_len * (4 + dtype.itemsize) / 1e12:.1f} TB",
                       "cache": "inputs larger than L2, no flush"},
            "formats": fm, "best_sell": best_sell,
            "roofline": {"bound": "hbm", "kernel": "csr", "achieved": fm["csr"]["gbs_rank0"], "peak": peak, "unit": "GB/s",
                         "frac": fm["csr"]["frac_measured_rank0"], "traffic": None, "peak_source": peak_src},
            "cpu_baseline": None, "e2e": None, "gpu_launches": int(args.steps * len(names)), "clocks": clk.summary(),
        }
        emit_json(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


def laplace_iter_arm(pkg, args, rank, world, local_rank):
    """BASELINE configs[4]: iterated SpMV (power iteration) on the 7-point Laplacian, fp64, rows
    partitioned over the ranks, NCCL all-gather of x once per step.  Weak scaling: every rank owns
    grid x grid x nz_per_gpu rows (400 x 400 x 50 = 8 M rows, 55.7 M nnz); 8 ranks = 400^3."""
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = pkg.Context(local_rank, stream=torch.cuda.current_stream().cuda_stream)
    L = pkg.lib()
    nx = ny = args.grid
    nz = args.nz_per_gpu * world
    n = nx * ny * nz
    blocks = pkg.equal_row_blocks(n, world, align=32)
    lo, hi = blocks.bounds(rank)
    n_local = hi - lo
    nnz = L.b200_gen_laplace7_nnz(nx, ny, nz, lo, n_local)
    rows, cols, vals = ctx.empty(nnz, np.int32), ctx.empty(nnz, np.int32), ctx.empty(nnz, np.float64)
    pkg.check(L.b200_gen_laplace7_coo(ctx.h, nx, ny, nz, lo, n_local, rows.ptr, cols.ptr, vals.ptr), "gen laplace7")
    pkg.check(L.b200_offset_i32(ctx.h, rows.ptr, nnz, -lo), "rebase rows")
    coo = pkg.CooMatrix(ctx, n_local, n, rows, cols, vals)
    fmt = args.iter_format
    csr = pkg.CsrMatrix(coo)
    mat = csr if fmt == "csr" else pkg.SellMatrix(csr, np.float64)
    if fmt == "csr":
        csr.plan()
    x_cur = torch.zeros(blocks.padded, dtype=torch.float64, device="cuda")
    x_next = torch.zeros_like(x_cur)
    pkg.check(L.b200_gen_uniform_f64(ctx.h, x_cur.data_ptr(), n, 11, 0.0, 1.0), "gen x0")
    calls = pkg.gpu_callables(pkg, ctx, mat, n_local)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    res = pkg.power_iteration(x_cur=x_cur, x_next=x_next, rank=rank, blocks=blocks, steps=args.warmup, **calls)
    x_cur, x_next = (res.x, x_next if res.x is x_cur else x_cur)
    sync_all()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        sync_all()
        t0.record()
        res = pkg.power_iteration(x_cur=x_cur, x_next=x_next, rank=rank, blocks=blocks, steps=args.steps, **calls)
        t1.record()
        sync_all()
    step_ms = t0.elapsed_time(t1) / args.steps

    # fused variant (SELL only): the SpMV kernel stores its y-block into every rank's next-x buffer
    # over NVLink (CUDA IPC peer pointers); one 1-element all-reduce per step is all that is left
    fused_ms = None
    fused_norm = None
    fused_full_ms = None
    fused_halo_ms = None
    ring_ms = None
    halo_bytes = None
    if fmt == "sell":
        bufs = pkg.PeerBuffers(pkg, ctx, blocks, rank, world)
        # halo-limited exchange (default): every rank learns which of its rows the others read as
        # columns (min/max column of each block, exchanged once) and the kernel stores only those
        ranges = pkg.exchange_col_ranges(pkg, ctx, coo.cols, lo, world)
        halo = pkg.halo_rows(ranges, blocks, rank)
        halo_bytes = 8 * sum(h - l for d, (l, h) in enumerate(zip(*halo)) if d != rank)

        def fused_run(h):
            pkg.check(L.b200_gen_uniform_f64(ctx.h, bufs.local[0].ptr, n, 11, 0.0, 1.0), "gen x0")
            pkg.check(L.b200_gen_uniform_f64(ctx.h, bufs.local[1].ptr, n, 11, 0.0, 1.0), "gen x0")
            sync_all()
            r0 = pkg.power_iteration_fused(pkg, ctx, mat, bufs, rank, blocks, args.warmup + (args.warmup % 2), halo=h)
            sync_all()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            r1 = pkg.power_iteration_fused(pkg, ctx, mat, bufs, rank, blocks, args.steps, first_step=r0.next_step,
                                           acc=r0.acc, halo=h)
            f1.record()
            sync_all()
            return f0.elapsed_time(f1) / args.steps, r1.norm

        fused_full_ms, _ = fused_run(None)
        fused_halo_ms, fused_halo_norm = fused_run(halo)

        # ring: no collective call in the loop; flags + partial sums travel through peer memory
        def ring_run():
            pkg.check(L.b200_gen_uniform_f64(ctx.h, bufs.local[bufs.ring_step % 2].ptr, n, 11, 0.0, 1.0), "gen x0")
            sync_all()
            pkg.power_iteration_ring(pkg, ctx, mat, bufs, rank, blocks, args.warmup + (args.warmup % 2), halo=halo)
            sync_all()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            r1 = pkg.power_iteration_ring(pkg, ctx, mat, bufs, rank, blocks, args.steps, halo=halo)
            f1.record()
            sync_all()
            return f0.elapsed_time(f1) / args.steps, r1.norm

        ring_ms, ring_norm = ring_run()
        fused_ms, fused_norm = fused_halo_ms, fused_halo_norm   # the product path
        bufs.close()

    # split: SpMV alone and the all-gather alone, same buffers (explains the step time)
    seg = x_next[rank * blocks.count:(rank + 1) * blocks.count]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        calls["spmv_local"](x_cur, seg)
    b.record()
    sync_all()
    spmv_ms = a.elapsed_time(b) / 10
    a.record()
    for _ in range(10):
        calls["all_gather_inplace"](x_next, seg)
    b.record()
    sync_all()
    gather_ms = a.elapsed_time(b) / 10

    t = torch.tensor([step_ms, spmv_ms, gather_ms, float(nnz), fused_ms or 0.0, fused_full_ms or 0.0,
                      float(halo_bytes or 0), ring_ms or 0.0], device="cuda", dtype=torch.float64)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        step_ms, spmv_ms, gather_ms, nnz_total = float(tmax[0]), float(tmax[1]), float(tmax[2]), float(tsum[3])
        fused_ms = float(tmax[4]) if fused_ms is not None else None
        fused_full_ms = float(tmax[5]) if fused_full_ms is not None else None
        halo_bytes = int(tmax[6]) if halo_bytes is not None else None
        ring_ms = float(tmax[7]) if ring_ms is not None else None
    else:
        nnz_total = float(nnz)
    nccl_ms = step_ms
    if fused_ms is not None:  # the product path when available; the NCCL formulation is reported next to it
        step_ms = fused_ms
    peak, peak_src = measured_peak()
    alg = mat.nbytes(np.float64)
    gbs = alg / (spmv_ms * 1e-3) * 1e-9
    if rank == 0:
        out = {
            "metric": "SpMV GFLOP/s inside the power iteration (2*nnz flops per step; step = SpMV + exchange of x "
                      "+ norm all-reduce)",
            "value": round(2.0 * nnz_total / (step_ms * 1e-3) * 1e-9, 2), "unit": "GFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(step_ms, 5),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"iterated SpMV (BASELINE configs[4]): 7-point Laplacian {nx}x{ny}x{nz} = {n} rows, "
                                   f"{int(nnz_total)} nnz, fp64, {fmt.upper()}, {world} row block(s), NCCL all-gather of x per step",
                       "rows_per_gpu": int(n_local), "nnz_per_gpu": int(nnz),
                       "cache": "inputs larger than L2 (0.7 GB matrix + 2 x 64 MB x per GPU per world rank), no flush"},
            "exchange": ("fused, halo-limited: the SpMV kernel stores each y row into the x buffers of the ranks that "
                         "read it (its own + neighbours) over NVLink (CUDA IPC) + one 256-byte NCCL all-reduce"
                         if fused_ms is not None else "NCCL all_gather_into_tensor (in place)"),
            "ring_no_collective": (None if ring_ms is None else
                                   {"ms_per_step": round(ring_ms, 5),
                                    "gflops": round(2.0 * nnz_total / (ring_ms * 1e-3) * 1e-9, 2),
                                    "what": "same halo-limited kernel + a one-warp kernel that hands over ||y||^2 partial "
                                            "sums and 'step done' flags through peer memory instead of the all-reduce "
                                            "(b200_spmv_sell_ring_f64); opt-in: measured slower than the all-reduce at "
                                            "2 GPUs"}),
            "fused_full_broadcast": (None if fused_full_ms is None else
                                     {"ms_per_step": round(fused_full_ms, 5),
                                      "gflops": round(2.0 * nnz_total / (fused_full_ms * 1e-3) * 1e-9, 2),
                                      "what": "same kernel, every row to every rank (what an all-gather moves)"}),
            "halo_bytes_sent_per_step_max_rank": halo_bytes,
            "nccl_allgather_formulation": {"ms_per_step": round(nccl_ms, 5),
                                           "gflops": round(2.0 * nnz_total / (nccl_ms * 1e-3) * 1e-9, 2),
                                           "split_ms": {"spmv": round(spmv_ms, 5), "all_gather": round(gather_ms, 5),
                                                        "other (sumsq, all-reduce, scale)":
                                                            round(max(nccl_ms - spmv_ms - gather_ms, 0.0), 5)},
                                           "eigenvalue_estimate": res.norm},
            "split_ms": {"spmv": round(spmv_ms, 5), "exchange + norm": round(max(step_ms - spmv_ms, 0.0), 5)},
            "eigenvalue_estimate": fused_norm if fused_ms is not None else res.norm,
            "roofline": {"bound": "hbm", "kernel": fmt, "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(gbs / peak, 4), "traffic": None, "peak_source": peak_src},
            "cpu_baseline": None, "e2e": None,
            "gpu_launches": int(args.steps * (1 if fused_ms is not None else 3)), "clocks": clk.summary(),
        }
        emit_json(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def reference_arm(pkg, args, dtype):
    """The reference's own CPU implementation on the host cores, same workload / metric / unit."""
    if args.workload == "banded":
        n_global = args.rows_per_gpu * max(args.gpus, 1)
        sr = min(args.cpu_sample_rows, args.rows_per_gpu)
        rows, cols, vals, x = banded_host(pkg, n_global, 0, sr)
        n_rows, n_cols = sr, n_global
        sample = (f"first {sr} rows ({len(rows)} nnz) of the banded matrix per step, five formats; fp64 because the "
                  "reference is fp64-only; compute_using_cpu of the unmodified coo/csr/ell/cmrs.c built -O3 "
                  "(oracle port for SELL: sigma_c.c has no CPU path)")
        workload = (f"banded FEM-like (BASELINE configs[2]): {args.rows_per_gpu} rows x {BANDED['nnz_per_row']} nnz/row "
                    f"per GPU; CPU sample = {sr} rows")
    else:
        n_rows, n_cols, rows, cols, vals, x = cant_host()
        sample = "the whole cant-shaped matrix per step, five formats, fp64"
        workload = "cant-shaped stand-in (BASELINE configs[1]): 62451 x 62451, 4325625 nnz, x = ramp"
    t0 = time.perf_counter()
    sec, detail, threads, kind, cdt = cpu_arm(n_rows, n_cols, rows, cols, vals, x, dtype, "reference",
                                              max(args.steps, 1), max(args.warmup, 0))
    value = 2.0 * len(rows) * len(FORMATS) / sec * 1e-9
    out = {
        "impl": "reference",
        "metric": "SpMV GFLOP/s, aggregate over the five formats (COO, CSR, ELL, SELL-32, CMRS); "
                  "per-format GFLOP/s and HBM GB/s (% of peak) in `formats`",
        "value": round(value, 3), "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": cdt, "data": "synthetic",
        "config": {"workload": workload, "formats": list(FORMATS)},
        "cpu_baseline": {"value": round(value, 3), "unit": "GFLOP/s", "cores": threads, "cpu_model": cpu_model(),
                         "omp_wait_policy": os.environ.get("OMP_WAIT_POLICY"), "kind": kind,
                         "sample": sample, "per_format_gflops": detail,
                         "as_shipped_no_O_flag": getattr(cpu_arm, "as_shipped", None),
                         "wall_s": round(time.perf_counter() - t0, 2)},
        "e2e": {"value": round(value, 3), "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_json(out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
