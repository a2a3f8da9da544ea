#!/usr/bin/env python
"""sweep_variants.py -- times the tuning variants of every SpMV kernel in ONE process on one GPU.

The tuning hooks (B200_CSR_LANES, B200_CSR_UNROLL, B200_ELL_LANES, B200_ELL_UNROLL, B200_SELL_WPC,
B200_SELL_UNROLL, B200_COO_U, B200_CMRS_U, B200_CMRS_WPS) are per-context options
(b200_ctx_set_option), so the sweep just sets them between launches.  Two workloads, both with cold
L2 data:

  cant    cant-shaped stand-in, 62 451 rows (fits in L2): N independent copies of every format used in
          rotation, the launches of one variant recorded into a CUDA graph and replayed;
  banded  2 097 152 rows x 64 nnz/row (1.1-1.6 GB per format): plain back-to-back launches.

    python opencl-spmv-algorithms_b200/tools/sweep_variants.py --workload cant --dtype f64 \
        --out gpurun_out/sweep_cant_f64.json

Measurement infrastructure only; the defaults it informs live in csrc/*.cu.
"""
from __future__ import annotations

import argparse
import itertools
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

HOOKS = ("B200_CMRS_STREAM", "B200_SELL_TMA", "B200_SELL_TMA_BLOCKS", "B200_CSR_LANES", "B200_CSR_UNROLL", "B200_ELL_LANES", "B200_ELL_UNROLL", "B200_SELL_WPC",
         "B200_SELL_UNROLL", "B200_COO_U", "B200_CMRS_U", "B200_CMRS_WPS", "B200_CSR_STREAM")


def variants(workload: str, args_no_tma: bool = False, tma_only: bool = False):
    small = workload == "cant"
    lanes = (2, 4, 8, 16) if small else (4, 8)
    out = {"coo": [{"B200_COO_U": u} for u in (1, 2, 4)],
           "cmrs": [{"B200_CMRS_U": u, "B200_CMRS_WPS": w} for u in (1, 2) for w in ((1, 2, 4) if small else (1,))],
           "cmrs_packed": [{"B200_CMRS_U": u, "B200_CMRS_WPS": w} for u in (1, 2) for w in ((1, 2, 4) if small else (1,))],
           "csr": [{"B200_CSR_LANES": l, "B200_CSR_UNROLL": u} for l, u in itertools.product(lanes, (1, 2, 4))],
           "ell": [{"B200_ELL_LANES": l, "B200_ELL_UNROLL": u} for l, u in itertools.product(lanes, (1, 2, 4))],
           "sell": [{"B200_SELL_WPC": w, "B200_SELL_UNROLL": u}
                    for w, u in itertools.product((1, 2, 4, 8) if small else (1,), (1, 2, 4))],
           "ell_colmajor": [{}]}
    if not args_no_tma:
        out["sell"] += [{"B200_SELL_TMA": 1, "B200_SELL_TMA_BLOCKS": b} for b in (1, 2, 3)]
    if small:
        out["csr"].append({"B200_CSR_STREAM": 1})
        for f in ("cmrs", "cmrs_packed"):   # the nnz-split kernel (atomics at run ends, like COO), forced
            out[f] += [{"B200_CMRS_STREAM": 1, "B200_CMRS_U": u} for u in (1, 2)]
    if tma_only:
        return {"sell": [{}] + [e for e in out["sell"] if "B200_SELL_TMA" in e]}
    for f in out:
        out[f].insert(0, {})  # the library's own default
    return out


CTX = None


def set_env(env):
    for k in HOOKS:
        CTX.set_option(k, None)
    for k, v in env.items():
        CTX.set_option(k, v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cant", choices=["cant", "banded"])
    ap.add_argument("--dtype", default="f64", choices=["f32", "f64"])
    ap.add_argument("--copies", type=int, default=7)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--rows", type=int, default=2097152)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-tma", action="store_true", help="skip the bulk-copy SELL variants")
    ap.add_argument("--tma-only", action="store_true", help="only the bulk-copy SELL variants (+ the default)")
    ap.add_argument("--overlap", action="store_true", help="b200_ctx_set_launch_overlap(1): PDL launches")
    ap.add_argument("--only", default=None, help="comma-separated formats to sweep (default: all)")
    args = ap.parse_args()
    dtype = np.dtype(np.float32 if args.dtype == "f32" else np.float64)

    import bench
    from __graft_entry__ import load_package
    pkg = load_package()
    ctx = pkg.Context(0)
    global CTX
    CTX = ctx
    peak, _ = bench.measured_peak()

    if args.workload == "cant":
        n_rows, n_cols, rows_h, cols_h, vals_h, x_h = bench.cant_host()
        x = ctx.array(x_h.astype(dtype))
        coos = [pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows_h, cols_h, vals_h) for _ in range(args.copies)]
    else:
        coo, x = bench.build_banded_device(pkg, ctx, args.rows, 0, args.rows, dtype)
        n_rows = args.rows
        coos = [coo]
    sets = []
    for c in coos:
        m = pkg.build_all(c, dtype)
        m["ell_colmajor"] = m.pop("ellcm")
        m["cmrs_packed"] = m["cmrs"].packed()
        sets.append(m)
    y = ctx.zeros(n_rows, dtype)
    ctx.set_l2_persist(x)
    ctx.sync()
    ctx.set_launch_overlap(args.overlap)
    nnz = coos[0].nnz
    n_launch = 4 * len(sets) if len(sets) > 1 else 10

    def reset_plans():
        # the CSR plan caches lanes-per-row / stream choice at creation: rebuild under the new env
        for st in sets:
            m = st["csr"]
            if m._plan:
                pkg.lib().b200_csr_plan_destroy(m._plan)
                m._plan = None
            m.plan()

    def reset_cmrs_plans():
        # ... and so does the CMRS plan (strip kernel vs nnz-split kernel); cmrs_packed shares it
        for st in sets:
            m = st["cmrs"]
            if getattr(m, "_plan", None):
                pkg.lib().b200_cmrs_plan_destroy(m._plan)
                m._plan = None
            m.plan()

    results = []
    for fmt, envs in variants(args.workload, args.no_tma, args.tma_only).items():
        if args.only and fmt not in args.only.split(","):
            continue
        nbytes = sets[0][fmt].nbytes(dtype)
        for env in envs:
            set_env(env)
            if fmt == "csr":
                reset_plans()
            if fmt in ("cmrs", "cmrs_packed"):
                reset_cmrs_plans()
            for st in sets:  # un-graphed pass: builds plans, loads the kernel
                st[fmt].spmv(x, y)
            ctx.sync()
            times = []
            if len(sets) > 1:
                with ctx.record_graph() as g:
                    for i in range(n_launch):
                        sets[i % len(sets)][fmt].spmv(x, y)
                g.launch()
                for _ in range(args.reps):
                    a, b = ctx.event(), ctx.event()
                    a.record()
                    g.launch()
                    b.record()
                    ctx.sync()
                    times.append(a.elapsed_ms_until(b) / n_launch)
            else:
                for _ in range(3):
                    sets[0][fmt].spmv(x, y)
                for _ in range(args.reps):
                    a, b = ctx.event(), ctx.event()
                    a.record()
                    for _ in range(n_launch):
                        sets[0][fmt].spmv(x, y)
                    b.record()
                    ctx.sync()
                    times.append(a.elapsed_ms_until(b) / n_launch)
            ms = float(np.median(times))
            rec = {"workload": args.workload, "dtype": args.dtype, "format": fmt, "env": env,
                   "ms": round(ms, 5), "ms_min": round(min(times), 5), "alg_bytes": int(nbytes),
                   "gbs": round(nbytes / (ms * 1e-3) * 1e-9, 1), "frac_measured": round(nbytes / (ms * 1e-3) * 1e-9 / peak, 4),
                   "gflops": round(2.0 * nnz / (ms * 1e-3) * 1e-9, 1)}
            results.append(rec)
            if args.out:
                Path(args.out).parent.mkdir(parents=True, exist_ok=True)
                Path(args.out).write_text(json.dumps(results, indent=1))
            print(f"{args.workload:6s} {args.dtype} {fmt:12s} {json.dumps(env):48s} {ms * 1e3:9.2f} us  "
                  f"{rec['gbs']:7.1f} GB/s  {rec['frac_measured']:.3f}", file=sys.stderr, flush=True)
    set_env({})
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(results, indent=1))
    ctx.close()


if __name__ == "__main__":
    main()
