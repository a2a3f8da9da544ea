#!/usr/bin/env python
"""ncu_target.py -- the smallest program that launches every SpMV kernel of one bench workload with the
library's DEFAULT variants: build the formats, then run each format's SpMV `--reps` times, nothing else.
Meant to sit behind `ncu --set full -k regex:...` (profiles/r2_scripts); it is also run plainly first.

    python opencl-spmv-algorithms_b200/tools/ncu_target.py --workload banded|cant|rmat|laplace [--dtype f32|f64]
"""
import argparse
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", required=True, choices=["banded", "cant", "rmat", "laplace"])
    ap.add_argument("--dtype", default=None)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--rmat-scale", type=int, default=24)
    ap.add_argument("--l2-persist", type=int, default=1)
    args = ap.parse_args()
    import bench
    from __graft_entry__ import load_package
    pkg = load_package()
    L = pkg.lib()
    ctx = pkg.Context(0)
    dname = args.dtype or {"banded": "f32", "cant": "f64", "rmat": "f32", "laplace": "f64"}[args.workload]
    dtype = np.dtype(np.float32 if dname == "f32" else np.float64)
    mats = {}
    if args.workload == "banded":
        coo, x = bench.build_banded_device(pkg, ctx, 2097152, 0, 2097152, dtype)
        m = pkg.build_all(coo, dtype)
        m["csr"].plan()
        mats = {f: m[f] for f in bench.FORMATS}
        mats["sell_delta16"] = pkg.Sell16Matrix(m["sell"])
        n_rows = 2097152
    elif args.workload == "cant":
        n_rows, n_cols, rows_h, cols_h, vals_h, x_h = bench.cant_host()
        coo = pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows_h, cols_h, vals_h)
        x = ctx.array(x_h.astype(dtype))
        m = pkg.build_all(coo, dtype)
        m["csr"].plan()
        mats = {f: m[f] for f in bench.FORMATS}
    elif args.workload == "laplace":
        nx = ny = 400
        nz = 50
        n_rows = nx * ny * nz
        nnz = L.b200_gen_laplace7_nnz(nx, ny, nz, 0, n_rows)
        rows, cols, vals = ctx.empty(nnz, np.int32), ctx.empty(nnz, np.int32), ctx.empty(nnz, np.float64)
        pkg.check(L.b200_gen_laplace7_coo(ctx.h, nx, ny, nz, 0, n_rows, rows.ptr, cols.ptr, vals.ptr), "gen")
        coo = pkg.CooMatrix(ctx, n_rows, n_rows, rows, cols, vals)
        x = ctx.empty(n_rows, dtype)
        pkg.check(L.b200_gen_uniform_f64(ctx.h, x.ptr, n_rows, 11, 0.0, 1.0), "x")
        csr = pkg.CsrMatrix(coo)
        csr.plan()
        sell = pkg.SellMatrix(csr, np.float64)
        mats = {"csr": csr, "sell": sell}
        # the fused kernel of the iterated mode, one destination (its own buffer)
        out = ctx.zeros(n_rows, np.float64)
        acc = ctx.zeros(32, np.float64)
        one = (C.c_void_p * 1)(out.ptr)

        class Fused:
            def spmv(self, xx, yy):
                pkg.check(L.b200_spmv_sell_halo_f64(ctx.h, sell.data.ptr, sell.cols.ptr, xx.ptr, sell.row_indices.ptr, 32,
                                                    sell.n_slices, n_rows, acc.ptr, acc.ptr, one, 1, 0, None, None), "fused")
        mats["sell_fused"] = Fused()
    else:
        scale, ef, abc, seed = args.rmat_scale, 16, (0.57, 0.19, 0.19), 5
        n_rows = 1 << scale
        cap = C.c_longlong(0)
        pkg.check(L.b200_gen_rmat_count(ctx.h, scale, ef, *abc, seed, 0, n_rows, C.byref(cap)), "count")
        rows, cols, vals = ctx.empty(cap.value, np.int32), ctx.empty(cap.value, np.int32), ctx.empty(cap.value, np.float64)
        nnz = C.c_longlong(0)
        pkg.check(L.b200_gen_rmat_coo(ctx.h, scale, ef, *abc, seed, 0, n_rows, cap.value, rows.ptr, cols.ptr, vals.ptr,
                                      C.byref(nnz)), "gen")
        rows.n = cols.n = vals.n = nnz.value
        coo = pkg.CooMatrix(ctx, n_rows, n_rows, rows, cols, vals)
        x = ctx.empty(n_rows, dtype)
        pkg.check((L.b200_gen_uniform_f32 if dname == "f32" else L.b200_gen_uniform_f64)(ctx.h, x.ptr, n_rows, 7, 0.0, 1.0), "x")
        csr = pkg.CsrMatrix(coo, check_sorted=False)
        csr.plan()
        mats = {"csr": csr, "coo": coo, "cmrs": pkg.CmrsMatrix(csr), "sell_sigma65536": pkg.SellMatrix(csr, dtype, sigma=65536, wide=True)}
    if args.l2_persist:
        ctx.set_l2_persist(x)
    y = ctx.zeros(n_rows, dtype)
    ctx.sync()
    for name, mat in mats.items():
        for _ in range(args.reps):
            mat.spmv(x, y)
            ctx.sync()     # one kernel at a time: no launch overlap under the profiler
    ctx.sync()
    print("ncu_target ok:", args.workload, dname, list(mats))
    ctx.close()


if __name__ == "__main__":
    main()
