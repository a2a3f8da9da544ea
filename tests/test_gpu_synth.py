"""GPU: device generators against their host twins, the iterated mode on one GPU against a CPU
power iteration, and -- at BASELINE.json's full banded size -- size-independent properties
(cross-format agreement, linearity, sampled rows against the oracle)."""
import numpy as np
import pytest

from __graft_entry__ import load_package
from oracle import binding as O
from test_distributed_cpu import laplace7

pytestmark = pytest.mark.gpu
pkg = load_package()


@pytest.fixture(scope="module")
def ctx():
    c = pkg.Context(0)
    yield c
    c.close()


@pytest.fixture(autouse=True)
def _reset_tuning_hooks(request):
    """Tuning hooks set with ctx.set_option (b200_ctx_set_option) never leak into the next test."""
    yield
    if "ctx" in request.fixturenames:
        request.getfixturevalue("ctx").clear_options()


def test_banded_generator_device_equals_host(ctx):
    L = pkg.lib()
    n, npr, hb, seed = 50000, 64, 2000, 42
    for r0, cnt in ((0, 3000), (20000, 4096), (47000, 3000)):
        nnz = cnt * npr
        rd, cd, vd = ctx.empty(nnz, np.int32), ctx.empty(nnz, np.int32), ctx.empty(nnz, np.float64)
        pkg.check(L.b200_gen_banded_coo(ctx.h, n, r0, cnt, npr, hb, seed, rd.ptr, cd.ptr, vd.ptr), "gen")
        rh, ch, vh = np.empty(nnz, np.int32), np.empty(nnz, np.int32), np.empty(nnz, np.float64)
        pkg.check(L.b200_gen_banded_coo_host(n, r0, cnt, npr, hb, seed, rh.ctypes.data, ch.ctypes.data,
                                             vh.ctypes.data), "gen host")
        assert np.array_equal(rd.download(), rh) and np.array_equal(cd.download(), ch)
        assert vd.download().tobytes() == vh.tobytes()
    xd = ctx.empty(100001, np.float64)
    pkg.check(L.b200_gen_uniform_f64(ctx.h, xd.ptr, 100001, 7, 0.0, 1.0), "gen x")
    xh = np.empty(100001)
    pkg.check(L.b200_gen_uniform_f64_host(xh.ctypes.data, 100001, 7, 0.0, 1.0), "gen x host")
    assert xd.download().tobytes() == xh.tobytes()
    assert 0.0 <= xh.min() and xh.max() < 1.0 and abs(xh.mean() - 0.5) < 0.01


def test_laplace7_generator(ctx):
    L = pkg.lib()
    nx, ny, nz = 13, 7, 9
    n, rows, cols, vals = laplace7(nx, ny, nz)
    assert L.b200_gen_laplace7_nnz(nx, ny, nz, 0, n) == rows.size == 7 * n - 2 * (nx * ny + ny * nz + nx * nz)
    for r0, cnt in ((0, n), (100, 333), (n - 50, 50)):
        sel = (rows >= r0) & (rows < r0 + cnt)
        nnz = L.b200_gen_laplace7_nnz(nx, ny, nz, r0, cnt)
        assert nnz == sel.sum()
        rd, cd, vd = ctx.empty(nnz, np.int32), ctx.empty(nnz, np.int32), ctx.empty(nnz, np.float64)
        pkg.check(L.b200_gen_laplace7_coo(ctx.h, nx, ny, nz, r0, cnt, rd.ptr, cd.ptr, vd.ptr), "gen")
        assert np.array_equal(rd.download(), rows[sel]) and np.array_equal(cd.download(), cols[sel])
        assert np.array_equal(vd.download(), vals[sel])


def test_power_iteration_one_gpu_matches_cpu(ctx):
    import torch
    nx, ny, nz, steps = 24, 20, 18, 40
    n, rows, cols, vals = laplace7(nx, ny, nz)
    blocks = pkg.equal_row_blocks(n, 1)
    x0 = np.zeros(blocks.padded)
    x0[:n] = np.random.default_rng(1).uniform(0, 1, n)
    # CPU power iteration with the oracle
    ptr, _ = O.build_csr(n, rows)
    x = x0[:n].copy()
    for _ in range(steps):
        y = O.spmv_csr(n, ptr, cols, vals, x)
        nrm = np.linalg.norm(y)
        x = y / nrm
    tctx = pkg.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    coo = pkg.CooMatrix.from_host(tctx, n, n, rows, cols, vals)
    csr = pkg.CsrMatrix(coo)
    for mat in (csr, pkg.SellMatrix(csr, np.float64)):
        x_cur = torch.from_numpy(x0.copy()).cuda()
        x_next = torch.zeros_like(x_cur)
        res = pkg.power_iteration(x_cur=x_cur, x_next=x_next, rank=0, blocks=blocks, steps=steps,
                                  **pkg.gpu_callables(pkg, tctx, mat, n))
        torch.cuda.synchronize()
        assert abs(res.norm - nrm) <= 1e-12 * nrm
        assert np.max(np.abs(res.x.cpu().numpy()[:n] - x)) <= 1e-12
    tctx.close()


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-5), (np.float64, 1e-12)])
def test_full_size_banded_properties(ctx, dtype, tol):
    """BASELINE configs[2] at full size (2 097 152 rows x 64 nnz/row; fp32 as benchmarked and fp64, the
    reference's arithmetic): too big for the CPU oracle in seconds, so check (a) all six kernels agree with
    each other, (b) linearity A(ax + bz) = aAx + bAz, (c) 8 x 1024 sampled rows -- first, last, around the
    middle and around every eighth -- against the oracle on the host twin of the generator."""
    L = pkg.lib()
    n, npr, hb, seed = 2097152, 64, 2000, 42
    nnz = n * npr
    rows, cols, vals = ctx.empty(nnz, np.int32), ctx.empty(nnz, np.int32), ctx.empty(nnz, np.float64)
    pkg.check(L.b200_gen_banded_coo(ctx.h, n, 0, n, npr, hb, seed, rows.ptr, cols.ptr, vals.ptr), "gen")
    coo = pkg.CooMatrix(ctx, n, n, rows, cols, vals)
    m = pkg.build_all(coo, dtype)
    assert m["csr"].plan_info().lanes_per_row == 4 and m["ell"].row_size == 64
    assert m["sell"].total == nnz and m["sell"].n_slices == n // 32
    rng = np.random.default_rng(0)
    xh, zh = rng.uniform(0, 1, n).astype(dtype), rng.uniform(-1, 1, n).astype(dtype)
    x, z = ctx.array(xh), ctx.array(zh)
    ys = {}
    for name, mat in m.items():
        yd = ctx.array(np.full(n, np.nan, dtype))
        mat.spmv(x, yd)
        ys[name] = yd.download().astype(np.float64)
    scale = np.abs(ys["csr"]).max()
    for name in ys:
        assert np.max(np.abs(ys[name] - ys["csr"])) / scale <= tol, name
    # linearity on the SELL kernel
    a, b = 0.75, -1.5
    w = ctx.array((a * xh + b * zh).astype(dtype))
    yz, yw = ctx.zeros(n, dtype), ctx.zeros(n, dtype)
    m["sell"].spmv(z, yz)
    m["sell"].spmv(w, yw)
    lin = a * ys["sell"] + b * yz.download().astype(np.float64)
    assert np.max(np.abs(yw.download() - lin)) / max(np.abs(lin).max(), 1e-30) <= tol
    # sampled row blocks against the oracle (host twin of the generator)
    for r0 in (0, 262144 - 100, 524288 + 7, 786432, 1048576 - 512, 1310720 + 33, 1835008 - 1000, n - 1024):
        cnt = 1024
        rh, ch, vh = np.empty(cnt * npr, np.int32), np.empty(cnt * npr, np.int32), np.empty(cnt * npr)
        pkg.check(L.b200_gen_banded_coo_host(n, r0, cnt, npr, hb, seed, rh.ctypes.data, ch.ctypes.data,
                                             vh.ctypes.data), "gen host")
        y_ref = O.yref(cnt, rh - r0, ch, vh, xh.astype(np.float64))
        for name in ys:
            assert O.rel_maxnorm(ys[name][r0:r0 + cnt], y_ref) <= tol, (name, r0)


def gen_rmat(ctx, scale, ef, r0, cnt, seed=3, abc=(0.57, 0.19, 0.19)):
    import ctypes as C
    L = pkg.lib()
    cand = C.c_longlong(0)
    pkg.check(L.b200_gen_rmat_count(ctx.h, scale, ef, *abc, seed, r0, cnt, C.byref(cand)), "rmat count")
    rows, cols, vals = ctx.empty(cand.value, np.int32), ctx.empty(cand.value, np.int32), ctx.empty(cand.value, np.float64)
    nnz = C.c_longlong(0)
    pkg.check(L.b200_gen_rmat_coo(ctx.h, scale, ef, *abc, seed, r0, cnt, cand.value, rows.ptr, cols.ptr,
                                  vals.ptr, C.byref(nnz)), "rmat gen")
    assert 0 < nnz.value <= cand.value
    return rows.download(nnz.value), cols.download(nnz.value), vals.download(nnz.value)


def test_rmat_generator_and_formats(ctx):
    """BASELINE configs[3] in small: the power-law generator (sorted, duplicate-free, a diagonal in
    every row, shard blocks concatenate to the whole matrix) and every kernel that is meant to run on
    it -- the nnz-split CSR kernel with rows spanning many tiles, COO, CMRS, SELL with a sigma sweep."""
    scale, ef = 13, 8
    n = 1 << scale
    rows, cols, vals = gen_rmat(ctx, scale, ef, 0, n)
    key = rows.astype(np.int64) << 32 | cols
    assert np.all(np.diff(key) > 0), "not sorted by (row, col) or duplicates left"
    assert cols.min() >= 0 and cols.max() < n and np.all(np.abs(vals) <= 1.0)
    assert set(np.flatnonzero(rows == cols)) and np.array_equal(rows[rows == cols], np.arange(n))
    lens = np.bincount(rows, minlength=n)
    assert lens.min() >= 1 and lens.max() > 20 * lens.mean()          # skewed
    parts = [gen_rmat(ctx, scale, ef, r0, cnt) for r0, cnt in ((0, 1000), (1000, 3096), (4096, n - 4096))]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), rows)
    assert np.array_equal(np.concatenate([p[1] for p in parts]), cols)
    assert np.concatenate([p[2] for p in parts]).tobytes() == vals.tobytes()

    x = np.random.default_rng(2).uniform(0, 1, n)
    y_ref = O.yref(n, rows, cols, vals, x)
    coo = pkg.CooMatrix.from_host(ctx, n, n, rows, cols, vals)
    csr = pkg.CsrMatrix(coo)
    info = csr.plan_info()
    assert info.stream_tiles > 0 and info.max_len > 512      # skewed: the nnz-split kernel is chosen
    padded = {}
    for dtype, tol in ((np.float64, 1e-12), (np.float32, 1e-5)):
        xd = ctx.array(x.astype(dtype))
        mats = {"coo": coo, "csr": csr, "cmrs": pkg.CmrsMatrix(csr)}
        for sigma in (1, 64, 4096, n):
            mats[f"sell{sigma}"] = pkg.SellMatrix(csr, dtype, sigma=sigma)
            padded[sigma] = mats[f"sell{sigma}"].total
        for name, mat in mats.items():
            yd = ctx.array(np.full(n, np.nan, dtype))
            mat.spmv(xd, yd)
            assert O.rel_maxnorm(yd.download(), y_ref) <= tol, (name, dtype)
        # hub chunks (> 256 columns) are split over many warps by the SELL plan; without the plan
        # one warp walks the whole chunk -- same answer
        assert mats["sell1"].plan_extra_items() > 0 and mats[f"sell{n}"].plan_extra_items() > 0
        # ... and so are CMRS strips longer than 8192 entries (none at this scale: forced in
        # test_cmrs_long_strips)
        yd = ctx.array(np.full(n, np.nan, dtype))
        mats["sell64"].spmv(xd, yd, use_plan=False)
        assert O.rel_maxnorm(yd.download(), y_ref) <= tol
    assert padded[1] > padded[64] > padded[4096] >= padded[n] >= rows.size
    perm, sp, _, _ = O.build_sell_sigma(n, rows, cols, vals, 4096)
    m = pkg.SellMatrix(csr, np.float64, sigma=4096)
    np.testing.assert_array_equal(m.perm.download(), perm)
    np.testing.assert_array_equal(m.slice_ptr.download(), sp)


def test_full_size_rmat_sampled_rows(ctx):
    """BASELINE configs[3] at FULL size (R-MAT scale 24: 16 777 216 rows, ~280 M nnz, fp32).  The CPU oracle
    cannot do this matrix in seconds, so: (a) sampled row blocks -- the block holding the longest (hub) row,
    the first rows, a block in the middle, the last rows -- are regenerated on their own (the generator is a
    pure function of the row block), multiplied by the oracle in fp64 and compared with every kernel's y;
    (b) all kernels agree with each other on ALL rows; (c) linearity of the CSR kernel."""
    import ctypes as C
    L = pkg.lib()
    scale, ef, abc, seed = 24, 16, (0.57, 0.19, 0.19), 5          # bench.py's matrix
    n = 1 << scale
    cap = C.c_longlong(0)
    pkg.check(L.b200_gen_rmat_count(ctx.h, scale, ef, *abc, seed, 0, n, C.byref(cap)), "count")
    rows, cols, vals = ctx.empty(cap.value, np.int32), ctx.empty(cap.value, np.int32), ctx.empty(cap.value, np.float64)
    nnz = C.c_longlong(0)
    pkg.check(L.b200_gen_rmat_coo(ctx.h, scale, ef, *abc, seed, 0, n, cap.value, rows.ptr, cols.ptr, vals.ptr,
                                  C.byref(nnz)), "gen")
    rows.n = cols.n = vals.n = nnz.value
    assert 2.5e8 < nnz.value < 3.0e8
    dtype = np.float32
    coo = pkg.CooMatrix(ctx, n, n, rows, cols, vals)
    csr = pkg.CsrMatrix(coo, check_sorted=False)
    info = csr.plan_info()
    assert info.stream_tiles > 0 and info.stream_tile_entries == 2048       # skewed: nnz-split, two groups per thread
    cmrs = pkg.CmrsMatrix(csr)
    assert cmrs.plan_stream_tiles() > 0                                      # ... and the nnz-split CMRS kernel
    mats = {"csr": csr, "coo": coo, "cmrs": cmrs, "cmrs_packed": cmrs.packed(),
            "sell_sigma65536": pkg.SellMatrix(csr, dtype, sigma=65536, wide=True)}
    rng = np.random.default_rng(0)
    xh, zh = rng.uniform(0, 1, n).astype(dtype), rng.uniform(-1, 1, n).astype(dtype)
    x, z = ctx.array(xh), ctx.array(zh)
    ys = {}
    for name, mat in mats.items():
        yd = ctx.array(np.full(n, np.nan, dtype))
        mat.spmv(x, yd)
        ys[name] = yd.download().astype(np.float64)
    scale_y = np.abs(ys["csr"]).max()
    for name in ys:
        assert np.max(np.abs(ys[name] - ys["csr"])) / scale_y <= 1e-5, name
    a, b = 0.75, -1.5
    w = ctx.array((a * xh + b * zh).astype(dtype))
    yz, yw = ctx.zeros(n, dtype), ctx.zeros(n, dtype)
    csr.spmv(z, yz)
    csr.spmv(w, yw)
    lin = a * ys["csr"] + b * yz.download().astype(np.float64)
    assert np.max(np.abs(yw.download() - lin)) / np.abs(lin).max() <= 1e-5
    ptr = csr.ptr.download()
    hub = int(np.argmax(np.diff(ptr)))
    assert ptr[hub + 1] - ptr[hub] > 100000                                  # a real hub row
    x64 = xh.astype(np.float64)
    for r0 in sorted({0, hub // 1024 * 1024, n // 2 + 4096, n - 1024}):
        rh, ch, vh = gen_rmat(ctx, scale, ef, r0, 1024, seed=seed, abc=abc)
        assert rh.size == ptr[r0 + 1024] - ptr[r0]                           # the block of the full build
        y_ref = O.yref(1024, rh - r0, ch, vh, x64)
        for name in ys:
            err = np.max(np.abs(ys[name][r0:r0 + 1024] - y_ref)) / scale_y
            assert err <= 1e-5, (name, r0, err)


def test_fused_power_iteration_one_gpu(ctx):
    """The fused SpMV + exchange kernel with a single destination (world = 1) must reproduce the
    NCCL-formulation iteration and the CPU power iteration."""
    import torch
    nx, ny, nz, steps = 24, 20, 18, 41
    n, rows, cols, vals = laplace7(nx, ny, nz)
    blocks = pkg.equal_row_blocks(n, 1)
    x0 = np.zeros(blocks.padded)
    x0[:n] = np.random.default_rng(1).uniform(0, 1, n)
    ptr, _ = O.build_csr(n, rows)
    x = x0[:n].copy()
    for _ in range(steps):
        y = O.spmv_csr(n, ptr, cols, vals, x)
        nrm = np.linalg.norm(y)
        x = y / nrm
    tctx = pkg.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    sell = pkg.SellMatrix(pkg.CsrMatrix(pkg.CooMatrix.from_host(tctx, n, n, rows, cols, vals)), np.float64)
    bufs = pkg.PeerBuffers(pkg, tctx, blocks, 0, 1)
    bufs.local[0].upload(x0)
    r1 = pkg.power_iteration_fused(pkg, tctx, sell, bufs, 0, blocks, 20)
    r2 = pkg.power_iteration_fused(pkg, tctx, sell, bufs, 0, blocks, steps - 20, first_step=r1.next_step, acc=r1.acc)
    tctx.sync()
    assert abs(r2.norm - nrm) <= 1e-12 * nrm
    y_un = r2.x.download()[:n]                      # unnormalised A x_{k-1}: normalise once at the end
    assert np.max(np.abs(y_un / np.linalg.norm(y_un) - x)) <= 1e-12
    bufs.close()
    tctx.close()


@pytest.mark.parametrize("mode,graph_steps", [("fused", 0), ("fused", 6), ("allgather", 0), ("allgather", 6)])
def test_library_iterator_one_gpu(ctx, mode, graph_steps):
    """b200_iterator_* (the C-ABI power iteration: steps issued by the library, optionally replayed
    from a launch graph) against the CPU power iteration, in both formulations, CSR and SELL blocks,
    and in pieces (7 + 30 + 4 steps: direct steps, graph replays, a direct remainder)."""
    nx, ny, nz, steps = 24, 20, 18, 41
    n, rows, cols, vals = laplace7(nx, ny, nz)
    blocks = pkg.equal_row_blocks(n, 1)
    x0 = np.zeros(blocks.padded)
    x0[:n] = np.random.default_rng(1).uniform(0, 1, n)
    ptr, _ = O.build_csr(n, rows)
    x = x0[:n].copy()
    for _ in range(steps):
        y = O.spmv_csr(n, ptr, cols, vals, x)
        nrm = np.linalg.norm(y)
        x = y / nrm
    csr = pkg.CsrMatrix(pkg.CooMatrix.from_host(ctx, n, n, rows, cols, vals))
    mats = [pkg.SellMatrix(csr, np.float64)] + ([csr] if mode == "allgather" else [])
    for mat in mats:
        bufs = [ctx.array(x0), ctx.zeros(blocks.padded, np.float64)]
        it = pkg.Iterator(pkg, ctx, None, mat, blocks, 0, 1, [[bufs[0].ptr], [bufs[1].ptr]], mode=mode,
                          graph_steps=graph_steps)
        for part in (7, 30, 4):
            it.run(part)
        norm = it.norm()
        k, xptr, launches = it.state()
        assert k == steps and xptr == bufs[steps % 2].ptr
        assert launches == steps * (1 if mode == "fused" else 3)
        assert abs(norm - nrm) <= 1e-12 * nrm
        got = bufs[steps % 2].download()[:n]
        if mode == "fused":
            got = got / norm          # the fused path leaves A x_{k-1} / ||x_{k-1}||, scaled by the next step
        assert np.max(np.abs(got - x)) <= 1e-12
        it.close()
    # argument checks: a communicator is required exactly when world > 1, graph_steps must be even
    with pytest.raises(pkg.B200Error):
        pkg.Iterator(pkg, ctx, None, mats[0], pkg.equal_row_blocks(n, 2), 0, 2, [[1, 1], [1, 1]], mode=mode)
    with pytest.raises(pkg.B200Error):
        pkg.Iterator(pkg, ctx, None, mats[0], blocks, 0, 1, [[bufs[0].ptr], [bufs[1].ptr]], mode=mode, graph_steps=3)


def test_iterated_80_cubed_matches_the_cpu_restatement(ctx):
    """the iterated mode on an 80^3 7-point Laplacian (512 000 rows) through the
    library's iterator (fused kernel, launch-graph replay) against the CPU restatement, same seeded x0 --
    the very check bench.py reports as iterated.oracle_80cubed."""
    import bench
    res = bench.iterated_oracle_check(pkg, ctx, g=80, steps=50)
    assert res["ok"] and res["rel_diff"] <= 1e-10, res
    assert 11.0 < res["norm_gpu"] < 12.0            # the Laplacian's largest eigenvalue is just below 12


@pytest.mark.parametrize("bcast_u", [None, 1, 2, 3, 4])
def test_halo_limited_exchange_three_emulated_ranks(ctx, bcast_u):
    """b200_spmv_sell_halo_f64: three row blocks of a 7-point Laplacian run one after the other on ONE
    GPU, each storing into the next-x buffers of all three 'ranks' but only the rows the destination
    reads (halo_rows from the blocks' column ranges).  The all-reduce of ||y||^2 is emulated on the
    host.  Must reproduce the CPU power iteration; rows outside a rank's block + halo must never be
    touched.  B200_BCAST_U: 1 = one chunk per warp, 2..4 = the persistent pipelined kernel at that many
    blocks per SM (the default picks it for large launches only, so it is forced here)."""
    import ctypes as C
    L = pkg.lib()
    if bcast_u is not None:
        ctx.set_option("B200_BCAST_U", bcast_u)
    nx, ny, nz, steps, G = 16, 12, 30, 25, 3
    n, rows, cols, vals = laplace7(nx, ny, nz)
    blocks = pkg.equal_row_blocks(n, G)
    rng = np.random.default_rng(5)
    x0 = np.zeros(blocks.padded)
    x0[:n] = rng.uniform(0, 1, n)
    ptr, _ = O.build_csr(n, rows)
    x = x0[:n].copy()
    for _ in range(steps):
        y = O.spmv_csr(n, ptr, cols, vals, x)
        nrm = np.linalg.norm(y)
        x = y / nrm
    sells, ranges, n_local = [], [], []
    for r in range(G):
        b0, b1 = blocks.bounds(r)
        sel = slice(ptr[b0], ptr[b1])
        coo = pkg.CooMatrix.from_host(ctx, b1 - b0, n, rows[sel] - b0, cols[sel], vals[sel])
        sells.append(pkg.SellMatrix(pkg.CsrMatrix(coo), np.float64))
        lo, hi = C.c_int(0), C.c_int(0)
        pkg.check(L.b200_minmax_i32(ctx.h, coo.cols.ptr, coo.nnz, C.byref(lo), C.byref(hi)), "minmax")
        assert (lo.value, hi.value) == (int(cols[sel].min()), int(cols[sel].max()))
        ranges.append((lo.value, hi.value))
        n_local.append(b1 - b0)
    halos = [pkg.halo_rows(ranges, blocks, r) for r in range(G)]
    plane = nx * ny
    # rank 1 reads rank 0's LAST plane, rank 2 nothing of it; rank 1 feeds its first plane to rank 0 and
    # its last plane to rank 2
    assert halos[0] == ([0, n_local[0] - plane, 0], [n_local[0], n_local[0], 0])
    assert halos[1] == ([0, 0, n_local[1] - plane], [plane, n_local[1], n_local[1]])
    SENTINEL = -7.0
    # bufs[r][b]: buffer b of emulated rank r; untouched entries keep the sentinel
    bufs = [[ctx.array(np.where(np.arange(blocks.padded) < n, x0, 0.0) if b == 0 else
                       np.full(blocks.padded, SENTINEL)) for b in range(2)] for r in range(G)]
    acc = [[ctx.zeros(32, np.float64) for _ in range(2)] for r in range(G)]
    for k in range(steps):
        cur, nxt = k % 2, (k + 1) % 2
        for r in range(G):
            dst = (C.c_void_p * G)(*[bufs[d][nxt].ptr for d in range(G)])
            lo_a, hi_a = (C.c_int * G)(*halos[r][0]), (C.c_int * G)(*halos[r][1])
            acc[r][k % 2].fill_bytes(0)
            pkg.check(L.b200_spmv_sell_halo_f64(
                ctx.h, sells[r].data.ptr, sells[r].cols.ptr, bufs[r][cur].ptr, sells[r].row_indices.ptr, 32,
                sells[r].n_slices, n_local[r], acc[r][(k - 1) % 2].ptr if k > 0 else None, acc[r][k % 2].ptr,
                dst, G, r * blocks.count, lo_a, hi_a), "b200_spmv_sell_halo_f64")
        total = np.zeros(32)
        for r in range(G):                      # the all-reduce
            total += acc[r][k % 2].download()
        for r in range(G):
            acc[r][k % 2].upload(total)
        ctx.sync()
    assert abs(np.sqrt(total.sum()) - nrm) <= 1e-12 * nrm
    last = steps % 2
    for r in range(G):
        got = bufs[r][last].download()
        b0, b1 = blocks.bounds(r)
        own = got[b0:b1]
        y_own = own / np.sqrt(total.sum())       # buffers hold the unnormalised A x_{k-1}
        assert np.max(np.abs(y_own - x[b0:b1])) <= 1e-12
        need_lo, need_hi = max(b0 - plane, 0), min(b1 + plane, n)
        outside = np.ones(blocks.padded, bool)
        outside[need_lo:need_hi] = False
        if steps % 2 == 1:                       # buffer 1 started as sentinels: nothing outside block + halo written
            assert np.all(got[outside] == SENTINEL), r
        assert not np.any(got[need_lo:need_hi] == SENTINEL)



@pytest.mark.parametrize("bcast_u", [1, 2, 3, 4])
def test_fused_kernel_variants_on_mixed_chunk_widths(ctx, bcast_u):
    """The fused kernel on a matrix whose SELL chunks are narrow (<= 8 columns: the pipelined lane = row
    path), wide (the 128-bit general path) and empty-rowed, in an order that alternates them inside one
    warp's walk, with a single destination: y / ||x|| and ||y||^2 against the oracle."""
    import ctypes as C
    L = pkg.lib()
    rng = np.random.default_rng(11)
    n = 32 * 700 + 13                                   # a ragged last chunk
    lens = rng.integers(0, 8, n)                         # narrow by default (some rows empty)
    for c0 in rng.choice(n // 32, 90, replace=False):    # ~13 % of the chunks get one long row
        lens[c0 * 32 + rng.integers(0, 32)] = rng.integers(9, 200)
    rows = np.repeat(np.arange(n, dtype=np.int32), lens)
    cols = rng.integers(0, n, rows.size).astype(np.int32)
    vals = rng.uniform(-1, 1, rows.size)
    x = rng.uniform(0.5, 1.5, n)
    prev = rng.uniform(0, 2, 32)                         # the 32 partial sums of ||x||^2 the kernel scales by
    y_ref = O.yref(n, rows, cols, vals, x) / np.sqrt(prev.sum())
    coo = pkg.CooMatrix.from_host(ctx, n, n, rows, cols, vals)
    sell = pkg.SellMatrix(pkg.CsrMatrix(coo, check_sorted=False), np.float64)
    ctx.set_option("B200_BCAST_U", bcast_u)
    xd, scale, acc = ctx.array(x), ctx.array(prev), ctx.zeros(32, np.float64)
    out = ctx.array(np.full(n + 64, -7.0))
    dst = (C.c_void_p * 1)(out.ptr)
    pkg.check(L.b200_spmv_sell_halo_f64(ctx.h, sell.data.ptr, sell.cols.ptr, xd.ptr, sell.row_indices.ptr, 32,
                                        sell.n_slices, n, scale.ptr, acc.ptr, dst, 1, 32, None, None), "fused kernel")
    got = out.download()
    assert np.all(got[:32] == -7.0) and np.all(got[32 + n:] == -7.0)       # offset respected, nothing past the end
    assert np.max(np.abs(got[32:32 + n] - y_ref)) <= 1e-12 * np.abs(y_ref).max()
    assert abs(acc.download().sum() - np.dot(y_ref, y_ref)) <= 1e-12 * np.dot(y_ref, y_ref)


def test_full_size_laplacian_all_rows(ctx):
    """BASELINE configs[4], one rank's block at full size (400 x 400 x 50 = 8 M rows, fp64): the device
    generator, the SELL build and the kernels the iterated mode runs at this size -- the persistent
    pipelined kernel behind the plain SELL SpMV and behind the fused SpMV + exchange call (1/||x|| scaling,
    ||y||^2, offset destination), and the CSR nnz-split kernel -- against the stencil applied with numpy
    slices to EVERY row."""
    import ctypes as C
    L = pkg.lib()
    nx = ny = 400
    nz = 50
    n = nx * ny * nz
    nnz = L.b200_gen_laplace7_nnz(nx, ny, nz, 0, n)
    assert nnz == 7 * n - 2 * (nx * ny + ny * nz + nx * nz)
    rows, cols, vals = ctx.empty(nnz, np.int32), ctx.empty(nnz, np.int32), ctx.empty(nnz, np.float64)
    pkg.check(L.b200_gen_laplace7_coo(ctx.h, nx, ny, nz, 0, n, rows.ptr, cols.ptr, vals.ptr), "gen")
    csr = pkg.CsrMatrix(pkg.CooMatrix(ctx, n, n, rows, cols, vals))
    sell = pkg.SellMatrix(csr, np.float64)
    assert sell.n_slices == n // 32 and sell.total <= 7 * n
    xh = np.random.default_rng(4).uniform(-1, 1, n)
    g = xh.reshape(nz, ny, nx)
    ref = 6.0 * g
    ref[1:, :, :] -= g[:-1, :, :]
    ref[:-1, :, :] -= g[1:, :, :]
    ref[:, 1:, :] -= g[:, :-1, :]
    ref[:, :-1, :] -= g[:, 1:, :]
    ref[:, :, 1:] -= g[:, :, :-1]
    ref[:, :, :-1] -= g[:, :, 1:]
    ref = ref.reshape(n)
    xd = ctx.array(xh)
    for name, mat in (("sell (pipelined)", sell), ("csr (nnz-split)", csr)):
        yd = ctx.array(np.full(n, np.nan))
        mat.spmv(xd, yd)
        assert O.rel_maxnorm(yd.download(), ref) <= 1e-12, name
    prev = np.random.default_rng(5).uniform(0, 2, 32)           # partial sums of ||x||^2 the kernel scales by
    scale, acc = ctx.array(prev), ctx.zeros(32, np.float64)
    out = ctx.array(np.full(n + 96, -7.0))
    dst = (C.c_void_p * 1)(out.ptr)
    pkg.check(L.b200_spmv_sell_halo_f64(ctx.h, sell.data.ptr, sell.cols.ptr, xd.ptr, sell.row_indices.ptr, 32,
                                        sell.n_slices, n, scale.ptr, acc.ptr, dst, 1, 64, None, None), "fused kernel")
    got = out.download()
    y_scaled = ref / np.sqrt(prev.sum())
    assert np.all(got[:64] == -7.0) and np.all(got[64 + n:] == -7.0)
    assert O.rel_maxnorm(got[64:64 + n], y_scaled) <= 1e-12
    assert abs(acc.download().sum() - np.dot(y_scaled, y_scaled)) <= 1e-12 * np.dot(y_scaled, y_scaled)
