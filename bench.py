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
(slowest) kernel; `cpu_baseline` is the oracle port on the host cores on a bounded sample; `e2e` is the
same step through the C ABI with x coming from and y going back to pinned host memory on every call
(the five calls of a step on 1, 2 or 5 in-order queues; the best is `value`).  The default line also
carries `f64` (the same matrix in the reference's arithmetic), `strong` (the fixed matrix cut into N
row blocks) and `iterated` (BASELINE configs[4]: power iteration on the row-partitioned Laplacian, the
fused SpMV + halo-exchange kernel with NCCL's all-reduce or the NVSwitch multicast hand-over, checked
against the SpMV + ncclAllGather formulation in the same run).

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
# (no OMP_PROC_BIND here: with it the OpenMP runtime pins the main thread to ONE core when torch
# initialises it, every thread created afterwards -- NCCL proxies, CUDA callback threads -- inherits that
# mask, and under torchrun all ranks end up on the same core: the iterated section ran 0.30 instead of
# 0.18 ms per step at 4 ranks, 0.55 instead of 0.19 at 8; profiles/r2_iter_probe_n4.json)

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
def bind_to_gpu_numa(local_rank: int) -> None:
    """Multi-rank runs: pin this process to the CPUs next to its GPU (NVML's ideal affinity) BEFORE any
    pinned host memory is allocated, so that every rank's host<->device copies stay on its own socket
    instead of all ranks sharing rank 0's.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


class Dist:
    """torch.distributed over NCCL as plumbing: barriers and the max / sum over ranks of the timings."""

    def __init__(self, world: int, local_rank: int):
        self.world, self.dist, self.torch = world, None, None
        if world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist, self.torch = dist, torch

    def barrier(self, ctx=None):
        if ctx is not None:
            ctx.sync()
        if self.dist is not None:
            self.torch.cuda.synchronize()
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def reduce(self, values, op="max"):
        if self.dist is None:
            return [float(v) for v in values]
        t = self.torch.tensor([float(v) for v in values], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(v) for v in t.cpu()]

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def column_segments(pkg, ctx, cols, n_cols, block_log2=12):
    """[(first column, count), ...]: the parts of x a row block reads, as runs of used 4096-column
    blocks (b200_used_column_blocks).  A banded shard reads one window of x -- two when the band wraps
    around the matrix edge -- so the host -> device bytes per rank do not grow with the number of ranks."""
    nb = (n_cols + (1 << block_log2) - 1) >> block_log2
    used = np.zeros(nb, np.uint8)
    pkg.check(pkg.lib().b200_used_column_blocks(ctx.h, cols.ptr, cols.n, n_cols, block_log2, used.ctypes.data),
              "b200_used_column_blocks")
    edges = np.flatnonzero(np.diff(np.concatenate(([0], used, [0]))))
    return [(int(s) << block_log2, min(int(e) << block_log2, n_cols) - (int(s) << block_log2))
            for s, e in zip(edges[0::2], edges[1::2])]


def measure_e2e(pkg, ctx, D, local_rank, mats, y, x, n_rows, n_cols, segments, dtype, steps, flops_step):
    """The metric through the C ABI with HOST buffers: per format, pinned-host x -> device, SpMV,
    y -> pinned host, with the format arrays uploaded once beforehand as the reference driver does
    (csr.c:183-193).  x lives in pinned host memory in full; only the column segments the row block
    reads (column_segments) are uploaded, so the bytes per rank do not grow with the number of ranks."""
    import ctypes as C
    L = pkg.lib()
    V = np.dtype(dtype).itemsize
    n_up = sum(c for _, c in segments)
    x_host = x.download()

    def make_queue(q):
        hx, hy = C.c_void_p(), C.c_void_p()
        pkg.check(L.b200_host_alloc_pinned(n_cols * V, C.byref(hx)), "pinned x")
        pkg.check(L.b200_host_alloc_pinned(n_rows * V, C.byref(hy)), "pinned y")
        np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_byte)), shape=(n_cols * V,)).view(dtype)[:] = x_host
        return dict(ctx=q, hx=hx, hy=hy, xin=q.zeros(n_cols, dtype))

    # queue 0 is the caller's context; the others are further in-order queues of the same device
    queues = [make_queue(ctx)]
    extra_ctx = [pkg.Context(local_rank) for _ in range(len(mats) - 1)]
    queues += [make_queue(c) for c in extra_ctx]

    def call(f, m, q):
        m.ctx = q["ctx"]
        for first, count in segments:
            pkg.check(L.b200_memcpy_h2d_async(q["ctx"].h, q["xin"].ptr + first * V, q["hx"].value + first * V, count * V), "h2d x")
        m.spmv(q["xin"], y[f])
        pkg.check(L.b200_memcpy_d2h_async(q["ctx"].h, q["hy"], y[f].ptr, n_rows * V), "d2h y")

    # The same five calls spread over K in-order queues of the same device (K = 1, 2, one per format): PCIe is
    # full duplex and the copy engines run next to the SMs, so the y download of one call overlaps the x upload
    # of the next and the kernel of a third.  Every call still moves its own x in and its own y out.
    csr_pinned = "csr" in mats and mats["csr"].plan_info().stream_tiles > 0  # a stream plan serves one queue only

    def owners(k):
        out, i = {}, 0
        for f in mats:
            if f == "csr" and csr_pinned:
                out[f] = queues[0]
                continue
            out[f] = queues[i % k]
            i += 1
        return out

    def timed(k):
        own = owners(k)
        used = queues[:k]

        def step():
            for f, m in mats.items():
                call(f, m, own[f])

        def drain():
            for q in used:
                q["ctx"].sync()
        for _ in range(2):
            step()
        drain()
        D.barrier(ctx)
        a, b = ctx.event(), ctx.event()
        t0 = time.perf_counter()
        a.record()
        for _ in range(steps):
            step()
        b.record()
        drain()
        D.barrier(ctx)
        wall = (time.perf_counter() - t0) * 1e3 / steps
        return wall, a.elapsed_ms_until(b) / steps

    one_wall, one_ms = timed(1)
    two_wall, _ = timed(2)
    all_wall, _ = timed(len(queues))
    for m in mats.values():
        m.ctx = ctx
    one_ms, one_wall, two_wall, all_wall = D.reduce([one_ms, one_wall, two_wall, all_wall], "max")
    h2d, d2h = D.reduce([len(mats) * n_up * V, len(mats) * n_rows * V], "sum")
    best = min(one_ms, two_wall, all_wall)
    for q in queues:
        q["ctx"].sync()
        L.b200_host_free_pinned(q["hx"])
        L.b200_host_free_pinned(q["hy"])
    del queues
    for c in extra_ctx:
        c.close()
    gf = lambda ms: round(flops_step / (ms * 1e-3) * 1e-9, 2)  # noqa: E731
    return {"value": gf(best), "unit": "GFLOP/s",
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": round(best, 4), "steps": steps,
            "x_columns_uploaded_per_rank": int(n_up), "x_columns_total": int(n_cols),
            "x_segments_rank0": [list(sg) for sg in segments[:8]],
            "one_queue": {"ms_per_step": round(one_ms, 4), "wall_ms_per_step": round(one_wall, 4), "gflops": gf(one_ms)},
            "two_queues": {"wall_ms_per_step": round(two_wall, 4), "gflops": gf(two_wall)},
            "queue_per_format": {"queues": len(mats), "wall_ms_per_step": round(all_wall, 4), "gflops": gf(all_wall)},
            "link_gbs_each_way": round(max(h2d, d2h) / (best * 1e-3) * 1e-9, 1),
            "what": "per format: x from pinned host memory -> device (only the 4096-column blocks this rank's row block "
                    "reads, b200_used_column_blocks), SpMV through the C ABI, y -> pinned host; format arrays uploaded "
                    "once before the timed region, as the reference driver does (csr.c:183-193). The five calls of a step "
                    "go to 1, 2 or 5 in-order queues (contexts) of the device, every call with its own x upload and y "
                    "download; value = the best of the three (one queue: CUDA events; several: wall clock, all queues "
                    "drained); max over ranks.  With one queue per format the kernels of different formats also overlap "
                    "each other's ramp and drain, so this number can come within a few percent of -- or just above -- the "
                    "device-resident `value`, whose five launches run strictly one after the other on one queue"}


def time_formats(ctx, D, mats, x, y, steps, warmup, local_rank=None):
    """W warm-up steps, then K timed steps (one SpMV per format and step): per-format CUDA events
    inside, whole-region events outside, barrier + sync on both sides."""
    for _ in range(warmup):
        for f in mats:
            mats[f].spmv(x, y[f])
    D.barrier(ctx)
    ev = {f: [(ctx.event(), ctx.event()) for _ in range(steps)] for f in mats}
    e0, e1 = ctx.event(), ctx.event()
    clk = ClockSampler(local_rank) if local_rank is not None else None
    if clk:
        clk.__enter__()
    D.barrier(ctx)
    e0.record()
    for s in range(steps):
        for f, m in mats.items():
            ev[f][s][0].record()
            m.spmv(x, y[f])
            ev[f][s][1].record()
    e1.record()
    D.barrier(ctx)
    if clk:
        clk.__exit__()
    total_ms = e0.elapsed_ms_until(e1)
    per_ms = {f: float(np.mean([a.elapsed_ms_until(b) for a, b in ev[f]])) for f in mats}
    return total_ms / steps, per_ms, clk


def format_table(names, per_ms, bytes_alg, nnz, peak):
    fm = {}
    for f in names:
        gbs = bytes_alg[f] / (per_ms[f] * 1e-3) * 1e-9
        fm[f] = {"ms": round(per_ms[f], 5), "gflops": round(2.0 * nnz / (per_ms[f] * 1e-3) * 1e-9, 2),
                 "alg_bytes": int(bytes_alg[f]), "gbs": round(gbs, 1),
                 "frac_measured": round(gbs / peak, 4), "frac_nominal_8TBs": round(gbs / NOMINAL_HBM_GBS, 4)}
    return fm


def strong_section(pkg, ctx, D, args, rank, world, dtype, peak):
    """STRONG scaling of the same five-format step: the fixed (--rows-per-gpu)-row banded matrix is cut
    into `world` nnz-balanced, 32-aligned row blocks by b200_partition_rows; x replicated, no
    collective.  The K steps are one launch-graph replay (at 8 ranks a kernel lasts ~25 us)."""
    n = args.rows_per_gpu
    npr = BANDED["nnz_per_row"]
    ptr_host = (np.arange(n + 1, dtype=np.int64) * npr).astype(np.int32)
    cuts = pkg.partition_rows(ptr_host, world, 32)
    r0, r1 = int(cuts[rank]), int(cuts[rank + 1])
    coo, x = build_banded_device(pkg, ctx, n, r0, r1 - r0, dtype)
    allm = pkg.build_all(coo, dtype)
    allm["csr"].plan()
    mats = {f: allm[f] for f in FORMATS}
    y = {f: ctx.zeros(r1 - r0, dtype) for f in mats}
    steps = max(10, min(args.steps, 100))
    for f in mats:
        mats[f].spmv(x, y[f])
    with ctx.record_graph() as g:
        for _ in range(steps):
            for f in mats:
                mats[f].spmv(x, y[f])
    g.launch()
    D.barrier(ctx)
    e0, e1 = ctx.event(), ctx.event()
    e0.record()
    g.launch()
    e1.record()
    D.barrier(ctx)
    (ms,) = D.reduce([e0.elapsed_ms_until(e1) / steps], "max")
    (nnz_total,) = D.reduce([coo.nnz], "sum")
    x_unread = (n - sum(c for _, c in column_segments(pkg, ctx, coo.cols, n))) * np.dtype(dtype).itemsize
    (bytes_max,) = D.reduce([sum(m.nbytes(dtype) - x_unread for m in mats.values())], "max")
    flops = 2.0 * nnz_total * len(mats)
    return {"value": round(flops / (ms * 1e-3) * 1e-9, 2), "unit": "GFLOP/s", "scaling": "strong",
            "ms_per_step": round(ms, 5), "steps": steps, "n_gpus": world,
            "workload": f"the fixed {n}-row banded matrix ({int(nnz_total)} nnz) cut into {world} row block(s) by "
                        "b200_partition_rows (nnz-balanced, cuts multiples of 32); x replicated; no collective",
            "cuts": [int(c) for c in cuts], "alg_bytes_per_step_max_rank": int(bytes_max),
            "gbs_max_rank": round(bytes_max / (ms * 1e-3) * 1e-9, 1),
            "frac_measured_max_rank": round(bytes_max / (ms * 1e-3) * 1e-9 / peak, 4),
            "method": "K steps recorded into one launch graph, one replay timed with CUDA events, max over ranks; "
                      "the five formats of a step evict each other from L2 (sum of arrays > L2 up to 8 ranks)"}


def iterated_section(pkg, D, args, rank, world, local_rank, steps, extras=False, oracle_check=True):
    """BASELINE configs[4]: power iteration on the 7-point Laplacian, fp64, rows partitioned over the
    ranks (weak scaling: grid x grid x nz_per_gpu rows per rank; 8 ranks of 400 x 400 x 50 = 400^3).
    Two formulations run for the SAME number of steps from the SAME x0 and must agree:
      fused      SELL kernel + halo-limited peer stores + one 256-byte all-reduce (the product path)
      allgather  CSR SpMV, sum of squares, all-reduce, scale, in-place ncclAllGather (what north_star names)
    Both are issued by the library (b200_iterator_*: launch graph of G steps, NCCL called from C)."""
    import ctypes as C
    L = pkg.lib()
    ctx = pkg.Context(local_rank)
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
    csr = pkg.CsrMatrix(coo)
    csr.plan()
    sell = pkg.SellMatrix(csr, np.float64)
    comm = pkg.Comm(pkg, ctx, rank, world) if world > 1 else None
    bufs = pkg.PeerBuffers(pkg, ctx, blocks, rank, world)
    ranges = pkg.exchange_col_ranges(pkg, ctx, coo.cols, lo, world)
    halo = pkg.halo_rows(ranges, blocks, rank)
    halo_bytes = 8 * sum(h - l for d, (l, h) in enumerate(zip(*halo)) if d != rank)
    # the K timed steps are ONE replay of a K-step launch graph: every replay of a graph that holds NCCL
    # nodes ends in a host callback (~0.5 ms on this box), so short graphs pay it every few steps
    K = max(10, min(steps, 200) // 2 * 2)
    G = K
    warm = 1 + 2 * G    # step 0 is always direct; two replays record + warm the graph of this parity

    def start_vector():
        ctx.sync()
        D.barrier(ctx)
        pkg.check(L.b200_gen_uniform_f64(ctx.h, bufs.local[0].ptr, n, 11, 0.0, 1.0), "gen x0")
        bufs.local[1].fill_bytes(0)
        D.barrier(ctx)

    def run(mode, matrix, h, graph_steps=G, mcast=None):
        start_vector()
        it = pkg.Iterator(pkg, ctx, comm, matrix, blocks, rank, world, bufs.ptrs[:2], mode=mode, halo=h,
                          graph_steps=graph_steps, mcast=mcast)
        it.run(warm)
        D.barrier(ctx)
        n0 = it.state()[2]
        e0, e1 = ctx.event(), ctx.event()
        e0.record()
        it.run(K)
        e1.record()
        D.barrier(ctx)
        ms = e0.elapsed_ms_until(e1) / K
        norm = it.norm()
        launches = it.state()[2] - n0
        it.close()
        return ms, norm, launches

    fused_ms, fused_norm, fused_launches = run("fused", sell, halo)
    # the same run once more with the NVML clock sampler on: a polling thread next to a 20-40 ms region of
    # lock-step collectives perturbs it (probe: 178 -> 185..400 us per step), so the headline is the run above
    with ClockSampler(local_rank) as clk:
        sampled_ms, _, _ = run("fused", sell, halo)
    ag_ms, ag_norm, _ = run("allgather", csr, None)
    direct_ms, direct_norm, _ = run("fused", sell, halo, graph_steps=0)   # same steps, launch by launch
    # the same kernel with the all-reduce + barrier done by the NVSwitch (multimem.red on a multicast block):
    # two launches per step and no NCCL call; where the box cannot set multicast up this stays null
    mc_ms = mc_norm = None
    mc_note = "one rank: nothing to exchange"
    if world > 1:
        sup = C.c_int(0)
        pkg.check(L.b200_mcast_supported(ctx.h, C.byref(sup)), "b200_mcast_supported")
        (all_sup,) = D.reduce([-float(sup.value)], "max")
        mc_note = "NVLink multicast not supported by the device / driver"
        if all_sup == -1.0:
            try:
                mc = pkg.McastBlock(pkg, ctx, rank, world)
                mc_ms, mc_norm, _ = run("fused_mcast", sell, halo, mcast=mc)
                ctx.sync()
                D.barrier(ctx)
                mc.close()
                mc_note = "ran"
            except pkg.B200Error as e:
                mc_note = f"multicast set-up refused: {e}"
    full_ms = None
    if extras:
        full_ms, _, _ = run("fused", sell, None)                           # every row to every rank

    # the SpMV kernel by itself (no exchange partner: one destination, its own buffer), same buffers
    start_vector()
    one = (C.c_void_p * 1)(bufs.local[1].ptr)
    acc = ctx.zeros(32, np.float64)

    def kernel_alone():
        pkg.check(L.b200_spmv_sell_halo_f64(ctx.h, sell.data.ptr, sell.cols.ptr, bufs.local[0].ptr, sell.row_indices.ptr,
                                            32, sell.n_slices, n_local, None, acc.ptr, one, 1, rank * blocks.count,
                                            None, None), "spmv alone")
    for _ in range(3):
        kernel_alone()
    a, b = ctx.event(), ctx.event()
    a.record()
    for _ in range(20):
        kernel_alone()
    b.record()
    ctx.sync()
    spmv_ms = a.elapsed_ms_until(b) / 20
    ycsr = ctx.zeros(n_local, np.float64)
    for _ in range(3):
        csr.spmv(bufs.local[0], ycsr)
    a.record()
    for _ in range(20):
        csr.spmv(bufs.local[0], ycsr)
    b.record()
    ctx.sync()
    csr_ms = a.elapsed_ms_until(b) / 20

    # e2e: this rank's block of x0 from pinned host memory, K steps, the norm and the block back to the host
    hx = C.c_void_p()
    pkg.check(L.b200_host_alloc_pinned(blocks.count * 8, C.byref(hx)), "pinned x block")
    start_vector()
    pkg.check(L.b200_memcpy_d2h(ctx.h, hx, bufs.local[0].ptr + 8 * rank * blocks.count, 8 * blocks.count), "x0 block")
    it = pkg.Iterator(pkg, ctx, comm, sell, blocks, rank, world, bufs.ptrs[:2], mode="fused", halo=halo, graph_steps=G)
    it.run(warm)
    D.barrier(ctx)
    t0 = time.perf_counter()
    pkg.check(L.b200_memcpy_h2d_async(ctx.h, it.state()[1] + 8 * rank * blocks.count, hx, 8 * blocks.count), "h2d x block")
    it.run(K)
    e2e_norm = it.norm()
    pkg.check(L.b200_memcpy_d2h(ctx.h, hx, it.state()[1] + 8 * rank * blocks.count, 8 * blocks.count), "d2h x block")
    D.barrier(ctx)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / K
    it.close()
    L.b200_host_free_pinned(hx)

    fused_ms, ag_ms, direct_ms, spmv_ms, csr_ms, e2e_ms, sampled_ms = D.reduce(
        [fused_ms, ag_ms, direct_ms, spmv_ms, csr_ms, e2e_ms, sampled_ms], "max")
    if mc_ms is not None:
        (mc_ms,) = D.reduce([mc_ms], "max")
    (nnz_total,) = D.reduce([nnz], "sum")
    (halo_max,) = D.reduce([halo_bytes], "max")
    if full_ms is not None:
        (full_ms,) = D.reduce([full_ms], "max")
    norms = D.reduce([fused_norm, ag_norm, direct_norm, -fused_norm, -ag_norm], "max")
    same_on_all_ranks = norms[0] == -norms[3] and norms[1] == -norms[4]
    rel = abs(fused_norm - ag_norm) / abs(ag_norm)
    rel_direct = abs(fused_norm - direct_norm) / abs(ag_norm)
    parity_ok = bool(rel <= 1e-10 and rel_direct <= 1e-10 and same_on_all_ranks)
    # The product path is the fused kernel with whichever hand-over is faster on this box AND reproduced the
    # all-gather formulation's norm: NVSwitch multicast where the box can set it up, else NCCL's all-reduce.
    nccl_ms = fused_ms
    mc_ok = mc_ms is not None and mc_norm is not None and abs(mc_norm - ag_norm) <= 1e-10 * abs(ag_norm)
    sync_used = "ncclAllReduce (256 bytes)"
    if mc_ok and mc_ms < fused_ms:
        fused_ms, sync_used = mc_ms, "NVSwitch multicast (b200_mcast_*: multimem.st + multimem.red, no NCCL call in the loop)"
    peak, peak_src = measured_peak()
    # x is counted over the rows this rank reads (its block + the halo planes), not the whole padded vector
    x_read = min(int(ranges[rank][1]) + 1, n) - max(int(ranges[rank][0]), 0) if world > 1 else n
    alg = sell.nbytes(np.float64) - 8 * (n - x_read)
    flops = 2.0 * nnz_total
    out = {
        "ms_per_step": round(fused_ms, 5), "value": round(flops / (fused_ms * 1e-3) * 1e-9, 2), "unit": "GFLOP/s",
        "steps": K, "warmup": warm, "n_gpus": world, "scaling": "weak",
        "workload": f"power iteration, 7-point Laplacian {nx}x{ny}x{nz} = {n} rows, {int(nnz_total)} nnz, fp64, "
                    f"{world} row block(s) of {n_local} rows",
        "exchange": "fused: SELL-32 kernel stores each y row into the x buffers of the ranks that read it (own block + "
                    f"halo planes, CUDA IPC peer memory) + one hand-over of ||y||^2 per step: {sync_used}; all issued by "
                    f"b200_iterator_run as ONE launch graph of {G} steps",
        "fused_with_nccl_allreduce": {"ms_per_step": round(nccl_ms, 5)},
        "halo_bytes_sent_per_step_max_rank": int(halo_max),
        "gpu_launches": int(fused_launches),
        "direct_launches_no_graph": {"ms_per_step": round(direct_ms, 5)},
        "ms_per_step_while_sampling_clocks": round(sampled_ms, 5),
        "nccl_allgather_formulation": {"ms_per_step": round(ag_ms, 5), "gflops": round(flops / (ag_ms * 1e-3) * 1e-9, 2),
                                       "what": "CSR SpMV into the rank's segment, sum of squares, 1-element all-reduce, "
                                               "scale, in-place ncclAllGather; same launch-graph replay"},
        "fused_full_broadcast": None if full_ms is None else {"ms_per_step": round(full_ms, 5)},
        "nvswitch_multicast": {"ms_per_step": None if mc_ms is None else round(mc_ms, 5), "norm": mc_norm, "status": mc_note,
                               "parity_ok": None if mc_norm is None else bool(abs(mc_norm - fused_norm) <= 1e-10 * abs(fused_norm)),
                               "what": "same fused kernel; the 32 partial sums are all-reduced and the ranks synchronised "
                                       "by multimem.red on a multicast block (b200_mcast_*): two launches per step, no NCCL"},
        "norm_fused": fused_norm, "norm_allgather": ag_norm, "norm_fused_direct": direct_norm,
        "rel_diff": rel, "parity_tol": 1e-10, "parity_ok": parity_ok,
        "parity": "same x0 (seeded), same step count in every run; |norm_fused - norm_allgather| <= 1e-10 * norm, the "
                  "graph replay equals the launch-by-launch run, every rank holds the same norms",
        "split_ms": {"spmv_kernel_alone": round(spmv_ms, 5), "exchange_and_norm": round(max(fused_ms - spmv_ms, 0.0), 5),
                     "exchange_and_norm_with_nccl": round(max(nccl_ms - spmv_ms, 0.0), 5),
                     "csr_stream_kernel_alone": round(csr_ms, 5)},
        "roofline": {"bound": "hbm", "kernel": "sell32 fused (alone)", "achieved": round(alg / (spmv_ms * 1e-3) * 1e-9, 1),
                     "peak": peak, "unit": "GB/s", "frac": round(alg / (spmv_ms * 1e-3) * 1e-9 / peak, 4),
                     "alg_bytes": int(alg), "traffic": ncu_traffic("laplace-iter", "f64", "sell_fused", n_local),
                     "peak_source": peak_src},
        "e2e": {"value": round(flops / (e2e_ms * 1e-3) * 1e-9, 2), "unit": "GFLOP/s", "ms_per_step": round(e2e_ms, 5),
                "h2d_bytes_per_step": int(8 * blocks.count * world / K), "d2h_bytes_per_step": int((8 * blocks.count + 8) * world / K),
                "norm": e2e_norm,
                "what": f"per rank: its x block from pinned host memory -> device, {K} steps, the norm and the block back "
                        "to pinned host memory; wall clock / steps, max over ranks"},
        "clocks": clk.summary(),
    }
    if not parity_ok:
        out["parity_failure"] = {"rel_fused_vs_allgather": rel, "rel_graph_vs_direct": rel_direct,
                                 "same_on_all_ranks": same_on_all_ranks}

    # N = 1: the same code on an 80^3 grid against the CPU restatement; the
    # restatement's timing is the section's cpu_baseline
    if oracle_check and world == 1 and rank == 0 and not args.no_cpu_baseline:
        out["oracle_80cubed"] = iterated_oracle_check(pkg, ctx)
    bufs.close()
    if comm:
        comm.close()
    ctx.close()
    return out


def iterated_oracle_check(pkg, ctx, g=80, steps=50):
    """Power iteration on a g^3 Laplacian: the library's fused iterator on the GPU vs the CPU
    restatement (oracle CSR SpMV + numpy norm), same x0 (the generators have bit-identical host twins)."""
    from oracle import binding as O
    L = pkg.lib()
    n = g * g * g
    nnz = L.b200_gen_laplace7_nnz(g, g, g, 0, n)
    rows, cols, vals = ctx.empty(nnz, np.int32), ctx.empty(nnz, np.int32), ctx.empty(nnz, np.float64)
    pkg.check(L.b200_gen_laplace7_coo(ctx.h, g, g, g, 0, n, rows.ptr, cols.ptr, vals.ptr), "gen laplace7")
    coo = pkg.CooMatrix(ctx, n, n, rows, cols, vals)
    sell = pkg.SellMatrix(pkg.CsrMatrix(coo), np.float64)
    blocks = pkg.equal_row_blocks(n, 1)
    x0 = np.zeros(blocks.padded)
    pkg.check(L.b200_gen_uniform_f64_host(x0.ctypes.data, n, 11, 0.0, 1.0), "gen x0 host")
    b = [ctx.array(x0), ctx.zeros(blocks.padded, np.float64)]
    it = pkg.Iterator(pkg, ctx, None, sell, blocks, 0, 1, [[b[0].ptr], [b[1].ptr]], mode="fused", graph_steps=10)
    it.run(steps)
    gpu_norm = it.norm()
    it.close()
    rh, ch, vh = rows.download(), cols.download(), vals.download()
    O.lib().orc_set_threads(os.cpu_count() or 1)
    ptr, _ = O.build_csr(n, rh)
    x = x0[:n].copy()
    t0 = time.perf_counter()
    for _ in range(steps):
        y = O.spmv_csr(n, ptr, ch, vh, x)
        cpu_norm = float(np.linalg.norm(y))
        x = y / cpu_norm
    sec = time.perf_counter() - t0
    rel = abs(gpu_norm - cpu_norm) / cpu_norm
    return {"grid": f"{g}^3", "steps": steps, "norm_gpu": gpu_norm, "norm_cpu_oracle": cpu_norm, "rel_diff": rel,
            "ok": bool(rel <= 1e-10),
            "cpu_baseline": {"value": round(2.0 * nnz * steps / sec * 1e-9, 3), "unit": "GFLOP/s",
                             "cores": os.cpu_count() or 1, "kind": "port",
                             "sample": f"{steps} power-iteration steps on the {g}^3 Laplacian ({nnz} nnz), oracle CSR SpMV "
                                       "(-O3, OpenMP) + numpy norm"}}


def ncu_traffic(workload, dname, kernel, n_rows):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture of this workload
    (profiles/ncu_traffic.json); None for workloads / sizes that were not captured."""
    try:
        cap = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
        for c in cap if isinstance(cap, list) else [cap]:
            if c["workload"] == workload and c["dtype"] == dname and c.get("rows", 2097152) == n_rows:
                return c["traffic_bytes"].get(kernel)
    except Exception:
        pass
    return None


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
    ap.add_argument("--rmat-sigmas", default="1,32,256,4096,65536,R")
    ap.add_argument("--l2-persist", type=int, default=None, choices=[0, 1],
                    help="pin x in L2 with an access-policy window (default: on for banded/cant, A/B reported for rmat)")
    ap.add_argument("--grid", type=int, default=400, help="laplace-iter / iterated: nx = ny")
    ap.add_argument("--nz-per-gpu", type=int, default=50, help="laplace-iter / iterated: z planes per rank")
    ap.add_argument("--iter-extras", action="store_true", help="laplace-iter: also time the full-broadcast variant")
    ap.add_argument("--dtype", default=None, choices=["f32", "f64"])
    ap.add_argument("--rows-per-gpu", type=int, default=2097152)
    ap.add_argument("--cant-copies", type=int, default=7,
                    help="cant: independent copies of the format arrays used in rotation (cold L2 without a flush)")
    ap.add_argument("--cpu-sample-rows", type=int, default=262144)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="banded: skip the fp64 pass and the `strong` / `iterated` sections (kernel A/B runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dtype = np.dtype(np.float32 if (args.dtype or ("f32" if args.workload == "banded" else "f64")) == "f32"
                     else np.float64)

    from __graft_entry__ import load_package
    pkg = load_package()

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(pkg, args, dtype)
    if world > 1:
        bind_to_gpu_numa(local_rank)
    if args.workload == "laplace-iter":
        return laplace_iter_arm(pkg, args, rank, world, local_rank)
    if args.workload == "rmat":
        return rmat_arm(pkg, args, rank, world, local_rank)
    return spmv_arm(pkg, args, rank, world, local_rank, dtype)


def spmv_arm(pkg, args, rank, world, local_rank, dtype):
    """The default line: one SpMV in each of the five formats per step (banded = BASELINE configs[2],
    weak scaling; cant = configs[1])."""
    dname = "f32" if dtype == np.float32 else "f64"
    D = Dist(world, local_rank)
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

    def build_set(coo_m, dt=dtype):
        allm = pkg.build_all(coo_m, dt)
        allm["csr"].plan()
        # "ell" = the kernel on the reference's row-major arrays (fastest ELL kernel on B200); the
        # column-major thread-per-row kernel is measured next to it
        # "cmrs_packed" = the derived 4+V bytes/entry layout (row_in_strip folded into the column word),
        # measured next to the reference two-array layout, never in the headline
        extras = {"ell_colmajor": allm["ellcm"], "cmrs_packed": allm["cmrs"].packed()}
        if args.workload == "banded":
            try:   # derived layout: 16-bit column deltas per SELL chunk (refused when a chunk spans > 65536
                   # columns); its kernel is the one-warp-per-chunk SELL kernel, i.e. for large matrices
                extras["sell_delta16"] = pkg.Sell16Matrix(allm["sell"])
            except pkg.B200Error:
                pass
        return ({f: allm[f] for f in FORMATS}, extras)

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
    # algorithmic bytes count x ONCE over the columns this row block actually reads (one or two windows of
    # the replicated x for a banded shard), not the whole replicated vector: at N ranks n_cols is N * R
    segments = column_segments(pkg, ctx, coo.cols, n_cols)
    x_unread = (n_cols - sum(c for _, c in segments)) * dtype.itemsize
    bytes_alg = {f: m.nbytes(dtype) - x_unread for f, m in {**mats, **extra}.items()}
    l2_persist = args.l2_persist != 0
    if l2_persist:
        ctx.set_l2_persist(x)
    ctx.sync()
    launch_no = [0]

    def next_set():
        launch_no[0] += 1
        return sets[launch_no[0] % n_copies]

    # old methodology, kept as context for the L2-sized workload only: a 256 MiB memset before every
    # launch (leaves the L2 full of DIRTY lines whose write-back competes with the kernel's reads)
    flush = ctx.empty(256 << 20, np.uint8) if args.workload == "cant" else None

    overlap = None
    if n_copies == 1:
        step_ms, per_ms, clk = time_formats(ctx, D, mats, x, y, args.steps, args.warmup, local_rank)
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

        def time_graph(g, reps, per):
            g.launch()
            out = []
            for _ in range(reps):
                a, b = ctx.event(), ctx.event()
                a.record()
                g.launch()
                b.record()
                ctx.sync()
                out.append(a.elapsed_ms_until(b) / per)
            return float(np.mean(out))

        for f in mats:  # un-graphed pass over every copy first: loads the kernels, warms the plans
            for st in sets:
                st[0][f].spmv(x, y[f])
        g_steps = record_steps(args.steps)
        g_warm = g_steps     # warm-up = one full replay of the K-step graph (>= W steps, same clocks / caches)
        n_f = 4 * n_copies
        g_fmt = {f: record_format(f, 0, n_f) for f in mats}
        e0, e1 = ctx.event(), ctx.event()
        with ClockSampler(local_rank) as clk:
            g_warm.launch()
            D.barrier(ctx)
            e0.record()
            g_steps.launch()
            e1.record()
            D.barrier(ctx)
            total_ms = e0.elapsed_ms_until(e1)
            per_ms = {f: time_graph(g_fmt[f], 5, n_f) for f in mats}
        step_ms = total_ms / args.steps

        # The numbers above are the library's default: back-to-back SpMV launches on the library's own
        # queue are programmatic dependents (every kernel streams its matrix arrays while its predecessor
        # drains and touches x / y only after it has finished; anything else entering the queue -- an upload,
        # a build, an event -- makes the next launch fully ordered again).  The same graphs recorded with
        # b200_ctx_set_launch_overlap(0), i.e. strictly one kernel after the other, are reported next to them.
        ctx.set_launch_overlap(False)
        g_steps_o = record_steps(args.steps)
        g_fmt_o = {f: record_format(f, 0, n_f) for f in mats}
        ctx.set_launch_overlap(True)
        g_steps_o.launch()
        D.barrier(ctx)
        e0.record()
        g_steps_o.launch()
        e1.record()
        D.barrier(ctx)
        overlap = {"ms_per_step": e0.elapsed_ms_until(e1) / args.steps,
                   "formats": {f: time_graph(g_fmt_o[f], 5, n_f) for f in mats}}

    # extra (untimed for the headline): the column-major ELL kernel and packed CMRS, same method
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
            per_ms[f] = time_graph(record_format(f, 1, n_f), 1, n_f)

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

    red = D.reduce([step_ms] + [per_ms[f] for f in mats], "max")
    step_ms = red[0]
    for i, f in enumerate(mats):
        per_ms[f] = red[1 + i]

    peak, peak_src = measured_peak()
    flops_step = 2.0 * nnz * len(mats) * world
    value = flops_step / (step_ms * 1e-3) * 1e-9
    fm = format_table(list(mats) + list(extra), per_ms, bytes_alg, nnz, peak)
    dom = max(mats, key=lambda f: per_ms[f])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": fm[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": fm[dom]["frac_measured"],
                "traffic": ncu_traffic(args.workload, dname, dom, n_rows) if world == 1 else None,
                "peak_source": peak_src, "per_format_frac": {f: fm[f]["frac_measured"] for f in mats}}

    # ---------------- e2e: host x in, host y out, through the C ABI, matrix resident ----------
    e2e = None
    steps_e = max(3, min(args.steps, 20))
    if not args.no_e2e:
        e2e = measure_e2e(pkg, ctx, D, local_rank, mats, y, x, n_rows, n_cols, segments, dtype, steps_e, flops_step)

    # ---------------- fp64 pass over the same matrix (the reference's own arithmetic) ----------
    f64 = None
    if args.workload == "banded" and dtype == np.float32 and not args.no_extras:
        del sets, extra
        for f in FORMATS:
            if f not in ("coo", "csr"):
                mats[f] = None       # release the fp32 ELL / SELL copies before building the fp64 ones
        d64 = np.dtype(np.float64)
        m64, _ = build_set(coo, d64)
        x64 = ctx.empty(n_cols, d64)
        pkg.check(L.b200_gen_uniform_f64(ctx.h, x64.ptr, n_cols, BANDED["x_seed"], 0.0, 1.0), "gen x")
        y64 = {f: ctx.zeros(n_rows, d64) for f in m64}
        if l2_persist:
            ctx.set_l2_persist(x64)
        k64 = max(10, min(args.steps, 50))
        s64, p64, _ = time_formats(ctx, D, m64, x64, y64, k64, args.warmup)
        red = D.reduce([s64] + [p64[f] for f in m64], "max")
        s64 = red[0]
        for i, f in enumerate(m64):
            p64[f] = red[1 + i]
        b64 = {f: m.nbytes(d64) - 2 * x_unread for f, m in m64.items()}   # x_unread counted 4-byte entries
        f64 = {"value": round(flops_step / (s64 * 1e-3) * 1e-9, 2), "unit": "GFLOP/s", "ms_per_step": round(s64, 5),
               "steps": k64, "dtype": "f64", "formats": format_table(list(m64), p64, b64, nnz, peak),
               "what": "the same matrix and step in fp64, the reference's only arithmetic (what --impl reference times)"}
        if not args.no_e2e:
            f64["e2e"] = measure_e2e(pkg, ctx, D, local_rank, m64, y64, x64, n_rows, n_cols, segments, d64,
                                     steps_e, flops_step)
        del m64, y64, x64
        ctx.set_l2_persist(None)

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

    # ---------------- the sections only a multi-GPU run can fill with meaning ------------------
    strong = iterated = None
    if args.workload == "banded" and not args.no_extras:
        del mats, y, coo, x
        strong = strong_section(pkg, ctx, D, args, rank, world, dtype, peak)
        ctx.sync()
        # the section times its OWN step count (reported as iterated.steps): one replay of a launch graph that holds
        # NCCL nodes ends in a ~0.5 ms host callback, which a 20-step graph would spread over 20 steps only
        iterated = iterated_section(pkg, D, args, rank, world, local_rank, 100)

    if rank == 0:
        out = {
            "metric": "SpMV GFLOP/s, aggregate over the five formats (COO, CSR, ELL, SELL-32, CMRS); "
                      "per-format GFLOP/s and HBM GB/s (% of peak) in `formats`",
            "value": round(value, 2), "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(step_ms, 5), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dname, "data": "synthetic",
            "config": {"workload": workload, "formats": list(FORMATS), "nnz_per_gpu": int(nnz),
                       "rows_per_gpu": int(n_rows), "cols": int(n_cols),
                       "partition": f"row blocks, {world} rank(s), x replicated, no collective",
                       "l2_persist_x": bool(l2_persist),
                       "cache": (f"inputs larger than L2 by rotation: {n_copies} independent copies of every format's "
                                 f"arrays ({min(bytes_alg.values()) >> 20}-{max(bytes_alg.values()) >> 20} MiB each), launch i "
                                 f"reads copy i mod {n_copies}; no flush; the K steps are one CUDA-graph replay" if n_copies > 1 else
                                 "inputs larger than L2 (0.8-1.6 GB per format), no flush")},
            "formats": fm,
            "launch_overlap_off": (None if overlap is None else {
                "what": "same K-step graph with b200_ctx_set_launch_overlap(0): strictly one kernel after the other "
                        "(the default lets back-to-back SpMV launches stream their matrix arrays while the previous one "
                        "drains; x is read / y written only after it has ended)",
                "ms_per_step": round(overlap["ms_per_step"], 5),
                "value": round(flops_step / (overlap["ms_per_step"] * 1e-3) * 1e-9, 2),
                "formats": {f: {"ms": round(ms, 5), "frac_measured": round(bytes_alg[f] / (ms * 1e-3) * 1e-9 / peak, 4)}
                            for f, ms in overlap["formats"].items()}}),
            "formats_flush_each_launch_context_only": flushed,
            "formats_warm_l2_context_only": warm, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "f64": f64, "strong": strong, "iterated": iterated,
            "gpu_launches": int(args.steps * len(FORMATS)), "clocks": clk.summary(),
        }
        emit_json(out)
    D.close()
    ctx.close()
    return 0


def rmat_arm(pkg, args, rank, world, local_rank):
    """BASELINE configs[3]: power-law R-MAT (scale 24: 16.8 M rows, ~270 M nnz after duplicate
    removal), fp32, CSR vs SELL-32-sigma sweep (+ COO, CMRS), rows partitioned nnz-balanced over the
    ranks (STRONG scaling: the global matrix is fixed), x replicated, no collective."""
    import ctypes as C
    D = Dist(world, local_rank)
    ctx = pkg.Context(local_rank)
    L = pkg.lib()
    scale, ef, abc, seed = args.rmat_scale, args.rmat_edge_factor, (0.57, 0.19, 0.19), 5
    n = 1 << scale
    dtype = np.dtype(np.float32 if (args.dtype or "f32") == "f32" else np.float64)
    V = dtype.itemsize

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

    def time_mat(m):
        for _ in range(args.warmup):
            m.spmv(x, y)
        D.barrier(ctx)
        a, b = ctx.event(), ctx.event()
        a.record()
        for _ in range(args.steps):
            m.spmv(x, y)
        b.record()
        D.barrier(ctx)
        return a.elapsed_ms_until(b) / args.steps

    peak, peak_src = measured_peak()
    results, order = {}, []

    def record(name, m, extra=None):
        ms = time_mat(m)
        alg = m.nbytes(dtype)
        results[name] = dict(ms=ms, alg_bytes=int(alg), nnz=int(nnz), **(extra or {}))
        order.append(name)

    # x (n * V bytes: 64 MB in fp32) fits the 126 MB L2: pin it with an access-policy window
    # (b200_ctx_set_l2_persist) unless --l2-persist 0; the A/B against the unpinned run is reported
    persist = args.l2_persist != 0
    cmrs = pkg.CmrsMatrix(csr)
    ab = {}
    with ClockSampler(local_rank) as clk:
        ctx.set_l2_persist(None)
        for name, m in (("csr", csr), ("coo", coo), ("cmrs", cmrs)):
            ab[name] = {"no_persist_ms": time_mat(m)}
        if persist:
            ctx.set_l2_persist(x)
        record("csr", csr)
        record("coo", coo)
        record("cmrs", cmrs)
        for name in ab:
            ab[name]["persist_ms" if persist else "no_persist_again_ms"] = results[name]["ms"]
        record("cmrs_packed", cmrs.packed())
        for tok in [t for t in args.rmat_sigmas.split(",") if t]:
            sigma = n_rows if tok == "R" else int(tok)
            total = C.c_longlong(0)
            sp = ctx.empty(L.b200_sell_num_slices(n_rows, 32) + 1, np.int64)
            perm = ctx.empty(n_rows, np.int32)
            pkg.check(L.b200_build_sell_ptr(ctx.h, csr.ptr.ptr, n_rows, 32, sigma, perm.ptr, sp.ptr,
                                            C.byref(total)), "sell ptr")
            name = f"sell_sigma{tok}"
            pad = total.value / max(nnz, 1)
            del sp, perm
            if total.value * (4 + V) > args.rmat_max_sell_bytes:
                results[name] = dict(ms=None, padding_factor=round(pad, 3), skipped="padded arrays exceed --rmat-max-sell-bytes")
                order.append(name)
                continue
            m = pkg.SellMatrix(csr, dtype, sigma=sigma, wide=True)
            record(name, m, dict(padding_factor=round(pad, 3)))
            del m
    # max over ranks per format, sum of nnz (strong scaling: the job is the whole matrix)
    names = [k for k in order if results[k].get("ms") is not None]
    ms = D.reduce([results[k]["ms"] for k in names], "max")
    (tot,) = D.reduce([float(nnz)], "sum")
    fm = {}
    for k in order:
        r = results[k]
        if r.get("ms") is None:
            fm[k] = r
            continue
        t_ms = float(ms[names.index(k)])
        gbs = r["alg_bytes"] / (r["ms"] * 1e-3) * 1e-9  # this rank's own bytes / own time
        fm[k] = {"ms": round(t_ms, 5), "gflops": round(2.0 * tot / (t_ms * 1e-3) * 1e-9, 2),
                 "alg_bytes_rank0": r["alg_bytes"], "gbs_rank0": round(gbs, 1), "frac_measured_rank0": round(gbs / peak, 4)}
        if "padding_factor" in r:
            fm[k]["padding_factor"] = r["padding_factor"]

    # e2e: host x in, host y out through the C ABI (CSR, COO, CMRS), matrix resident
    e2e = None
    if not args.no_e2e:
        mats3 = {"csr": csr, "coo": coo, "cmrs": cmrs}
        y3 = {f: y for f in mats3}
        e2e = measure_e2e(pkg, ctx, D, local_rank, mats3, y3, x, n_rows, n, column_segments(pkg, ctx, coo.cols, n), dtype,
                          max(3, min(args.steps, 10)), 2.0 * tot * len(mats3))
        e2e["formats"] = list(mats3)
        if world > 1:
            e2e["x_sharded_upload_nvlink_allgather"] = rmat_e2e_sharded(pkg, ctx, D, local_rank, rank, world, mats3, y, x, n_rows, n,
                                                                        dtype, max(3, min(args.steps, 10)), 2.0 * tot * len(mats3))

    # CPU baseline (rank 0, N = 1): the oracle port's CSR / COO / CMRS on the first rows of the same matrix
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import binding as O
        sr = min(1 << 20, n_rows)
        ptr_h = csr.ptr.download(sr + 1)
        k = int(ptr_h[sr])
        rh, ch, vh = rows.download(k), cols.download(k), vals.download(k)
        xh = x.download().astype(np.float64)
        threads = os.cpu_count() or 1
        O.lib().orc_set_threads(threads)
        sp_h, ris_h = O.build_cmrs(sr, rh)
        v_t, x_t = vh.astype(dtype), xh.astype(dtype)
        runs = {"csr": lambda: O.spmv_csr(sr, ptr_h, ch, v_t, x_t, dtype=dtype),
                "coo": lambda: O.spmv_coo(sr, rh, ch, v_t, x_t, dtype=dtype),
                "cmrs": lambda: O.spmv_cmrs(sr, sp_h, ris_h, ch, v_t, x_t, dtype=dtype)}
        per = {}
        for f, run in runs.items():
            run()
            best = float("inf")
            for _ in range(3):
                t0 = time.perf_counter()
                run()
                best = min(best, time.perf_counter() - t0)
            per[f] = best
        cpu = {"value": round(2.0 * k / per["csr"] * 1e-9, 3), "unit": "GFLOP/s", "cores": threads, "cpu_model": cpu_model(),
               "kind": "port", "dtype": "f32" if dtype == np.float32 else "f64",
               "sample": f"first {sr} rows ({k} nnz) of the same R-MAT matrix, oracle CSR (value), COO and CMRS, best of 3",
               "per_format_gflops": {f: round(2.0 * k / t * 1e-9, 3) for f, t in per.items()}}

    if rank == 0:
        best_sell = min((k for k in names if k.startswith("sell")), key=lambda k: fm[k]["ms"], default=None)
        out = {
            "metric": "SpMV GFLOP/s of CSR on the power-law matrix (2*nnz flops); COO, CMRS and the SELL-32-sigma "
                      "sweep with padding factors in `formats`",
            "value": fm["csr"]["gflops"], "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": fm["csr"]["ms"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if dtype == np.float32 else "f64", "data": "synthetic",
            "config": {"workload": f"R-MAT power-law (BASELINE configs[3]): scale {scale} = {n} rows, edge factor {ef}, "
                                   f"(a,b,c,d)=(0.57,0.19,0.19,0.05), diagonal added, duplicates removed: {int(tot)} nnz; "
                                   f"{world} nnz-balanced row block(s), x replicated, no collective",
                       "rank0": {"rows": int(n_rows), "nnz": int(nnz), "mean_len": round(info.mean_len, 2),
                                 "max_len": int(st.max_len), "csr_kernel": "nnz-split stream" if info.stream_tiles else
                                 f"vector, {info.lanes_per_row} lanes/row + {info.n_long_rows} long rows"},
                       "ell": f"not run: K = longest row = {int(st.max_len)} would need {n_rows * st.max_len * (4 + V) / 1e12:.1f} TB",
                       "l2_persist_x": bool(persist),
                       "cache": "inputs larger than L2, no flush"},
            "formats": fm, "best_sell": best_sell,
            "l2_persist_ab_rank0": {k: {kk: round(vv, 5) for kk, vv in v.items()} for k, v in ab.items()},
            "roofline": {"bound": "hbm", "kernel": "csr", "achieved": fm["csr"]["gbs_rank0"], "peak": peak, "unit": "GB/s",
                         "frac": fm["csr"]["frac_measured_rank0"],
                         "traffic": ncu_traffic("rmat", "f32" if dtype == np.float32 else "f64", "csr", n_rows) if world == 1 else None,
                         "peak_source": peak_src,
                         "note": "the x gather is random: this workload is bound by L1/L2 sector traffic of the gather, not "
                                 "by HBM (profiles/r2_ncu_rmat_summary.md)"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(args.steps * len(names)), "clocks": clk.summary(),
        }
        emit_json(out)
    D.close()
    ctx.close()
    return 0


def rmat_e2e_sharded(pkg, ctx, D, local_rank, rank, world, mats, y, x, n_rows, n_cols, dtype, steps, flops_step):
    """e2e of the row-partitioned single-shot SpMV with a REPLICATED x that comes from the host: every rank
    uploads only its 1/N slice of x, the ranks all-gather it over NVLink (b200_comm_allgather_bytes: NCCL called
    from the library), then SpMV, then this rank's rows of y back to pinned host memory.  Host-link bytes per call
    and rank: (n_cols / N + n_rows) * V instead of (n_cols + n_rows) * V."""
    import ctypes as C
    L = pkg.lib()
    V = np.dtype(dtype).itemsize
    per, padded = pkg.x_upload_slices(n_cols, world, V)      # 16-byte slices
    x_host = x.download()
    hx, hy = C.c_void_p(), C.c_void_p()
    pkg.check(L.b200_host_alloc_pinned(per * V, C.byref(hx)), "pinned x slice")
    pkg.check(L.b200_host_alloc_pinned(n_rows * V, C.byref(hy)), "pinned y")
    mine = np.zeros(per, dtype)
    lo = rank * per
    mine[:max(0, min(per, n_cols - lo))] = x_host[lo:lo + per]
    np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_byte)), shape=(per * V,)).view(dtype)[:] = mine
    xin = ctx.zeros(padded, dtype)
    comm = pkg.Comm(pkg, ctx, rank, world)

    def step():
        for m in mats.values():
            pkg.check(L.b200_memcpy_h2d_async(ctx.h, xin.ptr + lo * V, hx, per * V), "h2d x slice")
            comm.allgather_bytes(xin.ptr, per * V)
            m.spmv(xin, y)
            pkg.check(L.b200_memcpy_d2h_async(ctx.h, hy, y.ptr, n_rows * V), "d2h y")
    for _ in range(2):
        step()
    ctx.sync()
    ok = bool(np.array_equal(xin.download()[:n_cols], x_host))        # the gathered x is the host's x
    D.barrier(ctx)
    a, b = ctx.event(), ctx.event()
    a.record()
    for _ in range(steps):
        step()
    b.record()
    D.barrier(ctx)
    (ms,) = D.reduce([a.elapsed_ms_until(b) / steps], "max")
    (ok_all,) = D.reduce([0.0 if ok else 1.0], "max")
    h2d, d2h = D.reduce([len(mats) * per * V, len(mats) * n_rows * V], "sum")
    comm.close()
    L.b200_host_free_pinned(hx)
    L.b200_host_free_pinned(hy)
    return {"value": round(flops_step / (ms * 1e-3) * 1e-9, 2), "unit": "GFLOP/s", "ms_per_step": round(ms, 4), "steps": steps,
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "gathered_x_equals_host_x": ok_all == 0.0,
            "what": "per format: this rank's 1/N slice of x from pinned host memory -> device, all-gather of x over NVLink "
                    "(b200_comm_allgather_bytes), SpMV through the C ABI, y -> pinned host; one in-order queue, CUDA events, "
                    "max over ranks"}


def laplace_iter_arm(pkg, args, rank, world, local_rank):
    """BASELINE configs[4] as the headline: the `iterated` section of the default line, printed as the
    line itself (more steps, optionally the full-broadcast variant)."""
    D = Dist(world, local_rank)
    it = iterated_section(pkg, D, args, rank, world, local_rank, args.steps, extras=args.iter_extras)
    if rank == 0:
        oc = it.get("oracle_80cubed")
        out = {
            "metric": "SpMV GFLOP/s inside the power iteration (2*nnz flops per step; step = SpMV + exchange of x "
                      "+ norm all-reduce)",
            "value": it["value"], "unit": "GFLOP/s", "n_gpus": world, "steps": it["steps"], "warmup": it["warmup"],
            "ms_per_step": it["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "iterated SpMV (BASELINE configs[4]): " + it["workload"],
                       "cache": "inputs larger than L2 (0.7 GB matrix per GPU), no flush"},
            "iterated": it, "roofline": it["roofline"], "e2e": it["e2e"],
            "cpu_baseline": oc["cpu_baseline"] if oc else None,
            "gpu_launches": it["gpu_launches"], "clocks": it["clocks"],
        }
        emit_json(out)
    D.close()
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
