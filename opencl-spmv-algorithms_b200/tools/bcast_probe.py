#!/usr/bin/env python
"""bcast_probe.py -- A/B of the fused SpMV + exchange kernel variants (B200_BCAST_U) on one GPU.

    python opencl-spmv-algorithms_b200/tools/bcast_probe.py [--grid 400] [--nz 50] [--reps 50]

7-point Laplacian grid x grid x nz (BASELINE configs[4], one rank's block), fp64, SELL-32.  For every
variant: the kernel alone (CUDA events around `reps` launches), its algorithmic GB/s, and the library
iterator's fused step (memset + kernel, one launch graph of 100 steps).  Prints one JSON line.
Variants: 1 = one chunk per warp (sell32_bcast_kernel), 2 / 3 / 4 = persistent pipelined kernel at that
many blocks per SM, 0 = the library's choice.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=400)
    ap.add_argument("--nz", type=int, default=50)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--variants", default="1,2,3,4,0")
    args = ap.parse_args()
    pkg = load_package()
    L = pkg.lib()
    ctx = pkg.Context(0)
    nx = ny = args.grid
    n = nx * ny * args.nz
    nnz = L.b200_gen_laplace7_nnz(nx, ny, args.nz, 0, n)
    rows, cols, vals = ctx.empty(nnz, np.int32), ctx.empty(nnz, np.int32), ctx.empty(nnz, np.float64)
    pkg.check(L.b200_gen_laplace7_coo(ctx.h, nx, ny, args.nz, 0, n, rows.ptr, cols.ptr, vals.ptr), "gen")
    sell = pkg.SellMatrix(pkg.CsrMatrix(pkg.CooMatrix(ctx, n, n, rows, cols, vals)), np.float64)
    alg = sell.nbytes(np.float64)
    blocks = pkg.equal_row_blocks(n, 1)
    b = [ctx.zeros(blocks.padded, np.float64), ctx.zeros(blocks.padded, np.float64)]
    pkg.check(L.b200_gen_uniform_f64(ctx.h, b[0].ptr, n, 11, 0.0, 1.0), "x0")
    acc = ctx.zeros(32, np.float64)
    one = (C.c_void_p * 1)(b[1].ptr)
    out = {"rows": n, "nnz": int(nnz), "alg_bytes": int(alg), "variants": {}}
    ref = None
    for v in [int(t) for t in args.variants.split(",")]:
        ctx.set_option("B200_BCAST_U", None if v == 0 else v)

        def kernel():
            pkg.check(L.b200_spmv_sell_halo_f64(ctx.h, sell.data.ptr, sell.cols.ptr, b[0].ptr, sell.row_indices.ptr, 32,
                                                sell.n_slices, n, None, acc.ptr, one, 1, 0, None, None), "kernel")
        for _ in range(5):
            kernel()
        e0, e1 = ctx.event(), ctx.event()
        e0.record()
        for _ in range(args.reps):
            kernel()
        e1.record()
        ctx.sync()
        ms = e0.elapsed_ms_until(e1) / args.reps
        y = b[1].download()[:n]
        if ref is None:
            ref = y
        # the library iterator: memset + kernel per step, 100 steps as one launch graph
        pkg.check(L.b200_gen_uniform_f64(ctx.h, b[0].ptr, n, 11, 0.0, 1.0), "x0")
        it = pkg.Iterator(pkg, ctx, None, sell, blocks, 0, 1, [[b[0].ptr], [b[1].ptr]], mode="fused", graph_steps=100)
        it.run(201)
        ctx.sync()
        e0.record()
        it.run(100)
        e1.record()
        ctx.sync()
        step_ms = e0.elapsed_ms_until(e1) / 100
        norm = it.norm()
        it.close()
        pkg.check(L.b200_gen_uniform_f64(ctx.h, b[0].ptr, n, 11, 0.0, 1.0), "x0")
        out["variants"][str(v)] = {"kernel_ms": round(ms, 5), "gbs": round(alg / (ms * 1e-3) * 1e-9, 1),
                                   "iterator_step_ms": round(step_ms, 5), "norm_after_301_steps": norm,
                                   "max_abs_diff_vs_first": float(np.max(np.abs(y - ref)))}
    # the plain SpMV (b200_spmv_sell_*, plan attached) of the same matrix: B200_SELL_PIPE = 0 (one chunk per
    # warp) | 2 | 3 | 4 (persistent pipelined kernel at that many blocks per SM) | unset (the plan's choice)
    ctx.set_option("B200_BCAST_U", None)
    out["plain_spmv"] = {}
    csr = sell.csr
    for dt in (np.float64, np.float32):
        m = sell if dt == np.float64 else pkg.SellMatrix(csr, dt)
        xs, ys = ctx.zeros(n, dt), ctx.zeros(n, dt)
        gen = L.b200_gen_uniform_f64 if dt == np.float64 else L.b200_gen_uniform_f32
        pkg.check(gen(ctx.h, xs.ptr, n, 11, 0.0, 1.0), "x")
        alg_p = m.nbytes(dt)
        res, first = {}, None
        for v in (0, 2, 3, 4, None):
            ctx.set_option("B200_SELL_PIPE", v)
            for _ in range(5):
                m.spmv(xs, ys)
            e0, e1 = ctx.event(), ctx.event()
            e0.record()
            for _ in range(args.reps):
                m.spmv(xs, ys)
            e1.record()
            ctx.sync()
            ms = e0.elapsed_ms_until(e1) / args.reps
            got = ys.download()
            first = got if first is None else first
            res["auto" if v is None else str(v)] = {"ms": round(ms, 5), "gbs": round(alg_p / (ms * 1e-3) * 1e-9, 1),
                                                    "bit_identical_to_pipe0": bool(np.array_equal(got, first))}
        out["plain_spmv"]["f64" if dt == np.float64 else "f32"] = {"alg_bytes": int(alg_p), "variants": res}
    ctx.set_option("B200_SELL_PIPE", None)
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
