"""GPU parity tests proper: every call goes through the C ABI (libb200spmv.so) and is compared with
the CPU oracle on the same seeded inputs.  Integer/index arrays: bit-exact.  y: relative max-norm
<= 1e-12 (fp64) / 1e-5 (fp32) against the serial fp64 COO reference (check_result semantics,
inc/helper_functions.h:207-219) -- the tolerances BASELINE.json states."""
import ctypes as C

import numpy as np
import pytest

from __graft_entry__ import load_package
from conftest import random_sorted_matrix
from oracle import binding as O

pytestmark = pytest.mark.gpu
pkg = load_package()
TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


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


def check_y(name, y, y_ref, dtype):
    err = O.rel_maxnorm(y, y_ref)
    assert err <= TOL[np.dtype(dtype)], f"{name} {np.dtype(dtype).name}: rel max-norm {err:g}"


def run_all_formats(ctx, n_rows, n_cols, rows, cols, vals, dtype, x=None, coo_order=None,
                    on_domain=True):
    """Build the five formats on the GPU, compare every array with the oracle, run every kernel."""
    x = np.arange(n_cols, dtype=np.float64) if x is None else x   # csr.c:95-99
    y_ref = O.yref(n_rows, rows, cols, vals, x)
    coo = pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals)
    coo_any = None
    if coo_order is not None:
        coo_any = pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows[coo_order], cols[coo_order],
                                          vals[coo_order])
    m = pkg.build_all(coo, dtype, coo_any=coo_any)
    xd = ctx.array(x.astype(dtype))
    lens = np.bincount(rows, minlength=n_rows)

    # ---- builds: bit-exact against the oracle (which is pinned to the reference) ----
    ptr_ref = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    np.testing.assert_array_equal(m["csr"].ptr.download(), ptr_ref)
    if on_domain:
        ptr_o, changes = O.build_csr(n_rows, rows)
        assert changes == n_rows - 1
        np.testing.assert_array_equal(ptr_o, ptr_ref)
        K = int(lens.max())
        hi, lo, tot, last = O.ell_stats(n_rows, rows)
        st = m["csr"].row_stats()
        assert (st.max_len, st.min_len, st.sum_len) == (K, int(lens.min()), int(lens.sum()))
        if n_rows > 1:
            assert (st.max_len_excl_last, st.min_len_excl_last, st.sum_len_excl_last, st.last_len) \
                == (hi, lo, tot, last)
        assert m["ell"].row_size == K
        if last <= hi and n_rows > 1:  # the reference's well-defined ELL domain (quirk q3)
            ec, ed = O.build_ell(n_rows, hi, rows, cols, vals)
            np.testing.assert_array_equal(m["ell"].cols.download(), ec)
            assert m["ell"].data.download().tobytes() == ed.astype(dtype).tobytes()
            cc, cd = O.ell_to_colmajor(n_rows, hi, m["ellcm"].pitch, ec, ed)
            np.testing.assert_array_equal(m["ellcm"].cols.download(), cc)
            assert m["ellcm"].data.download().tobytes() == cd.astype(dtype).tobytes()
        ri, sc, sd = O.build_sell(n_rows, rows, cols, vals)
        np.testing.assert_array_equal(m["sell"].row_indices.download(), ri)
        np.testing.assert_array_equal(m["sell"].slice_ptr.download(), ri.astype(np.int64))
        np.testing.assert_array_equal(m["sell"].cols.download(), sc)
        assert m["sell"].data.download().tobytes() == sd.astype(dtype).tobytes()
        sp, ris = O.build_cmrs(n_rows, rows)
        np.testing.assert_array_equal(m["cmrs"].strip_ptr.download(), sp)
        np.testing.assert_array_equal(m["cmrs"].row_in_strip.download(), ris)

    # ---- SpMV: every kernel, y poisoned first so unwritten rows are caught ----
    for name, mat in m.items():
        yd = ctx.array(np.full(n_rows, np.nan, dtype))
        mat.spmv(xd, yd)
        check_y(name, yd.download(), y_ref, dtype)
    # CSR without a plan (analysed on the fly) gives the same bits as with one
    y1, y2 = ctx.zeros(n_rows, dtype), ctx.zeros(n_rows, dtype)
    m["csr"].spmv(xd, y1, use_plan=True)
    m["csr"].spmv(xd, y2, use_plan=False)
    assert y1.download().tobytes() == y2.download().tobytes()
    # SELL writing the reference's padded output (n_slices*32, sigma_c.c:212): padding rows = 0
    ns = m["sell"].n_slices
    yp = ctx.array(np.full(ns * 32, np.nan, dtype))
    m["sell"].spmv(xd, yp, n_out=ns * 32)
    got = yp.download()
    check_y("sell-padded", got[:n_rows], y_ref, dtype)
    assert np.all(got[n_rows:] == 0)
    return m, y_ref


SHAPES = [  # n_rows, n_cols, min_len, max_len, seed, long_rows
    (1, 40, 3, 3, 0, ()),
    (2, 9, 1, 4, 1, ()),
    (31, 64, 1, 9, 2, ()),
    (32, 64, 1, 9, 3, ()),
    (33, 64, 1, 9, 4, ()),
    (257, 300, 1, 70, 5, ()),
    (1000, 1500, 1, 3, 6, ()),
    (1000, 5000, 20, 120, 7, ()),
    (4099, 6000, 1, 40, 8, ((0, 40), (4098, 1))),
    (3000, 60000, 1, 12, 9, ((5, 50000), (1777, 9000), (2999, 700))),   # long-row kernel, splits
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", SHAPES, ids=[f"r{s[0]}-l{s[2]}-{s[3]}" for s in SHAPES])
def test_formats_random(ctx, shape, dtype):
    n_rows, n_cols, lo, hi, seed, long_rows = shape
    rows, cols, vals = random_sorted_matrix(n_rows, n_cols, lo, hi, seed, long_rows)
    order = np.lexsort((rows, cols))  # column-major file order, as cant.mtx for coo.c
    run_all_formats(ctx, n_rows, n_cols, rows, cols, vals, dtype, coo_order=order)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_random_x_and_unsorted_coo(ctx, dtype):
    rows, cols, vals = random_sorted_matrix(2000, 2000, 1, 60, 11)
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, 2000)
    run_all_formats(ctx, 2000, 2000, rows, cols, vals, dtype, x=x, coo_order=rng.permutation(rows.size))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_empty_rows_are_handled(ctx, dtype):
    """Off the reference's domain (quirk q4: its builders shift everything on an empty row); the
    GPU builders give the mathematically right arrays."""
    rows, cols, vals = random_sorted_matrix(500, 700, 1, 20, 12)
    keep = ~np.isin(rows, [0, 17, 18, 19, 255, 499])
    rows, cols, vals = rows[keep], cols[keep], vals[keep]
    m, y_ref = run_all_formats(ctx, 500, 700, rows, cols, vals, dtype, on_domain=False)
    assert y_ref[17] == 0 and y_ref[499] == 0


def test_degenerate_sizes(ctx):
    for dtype in (np.float64, np.float32):
        # no rows at all
        coo = pkg.CooMatrix.from_host(ctx, 0, 5, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0))
        csr = pkg.CsrMatrix(coo)
        assert list(csr.ptr.download()) == [0]
        x, y = ctx.array(np.ones(5, dtype)), ctx.zeros(1, dtype)
        csr.spmv(x, y)
        coo.spmv(x, y)
        # rows but no entries
        coo = pkg.CooMatrix.from_host(ctx, 70, 5, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0))
        m = pkg.build_all(coo, dtype)
        for name, mat in m.items():
            yd = ctx.array(np.full(70, np.nan, dtype))
            mat.spmv(x, yd)
            assert np.all(yd.download() == 0), name


def test_unsorted_rows_rejected(ctx):
    coo = pkg.CooMatrix.from_host(ctx, 4, 4, np.array([0, 2, 1, 3], np.int32),
                                  np.array([0, 1, 2, 3], np.int32), np.ones(4))
    with pytest.raises(pkg.B200Error) as e:
        pkg.CsrMatrix(coo)
    assert e.value.status == pkg.ERR_DOMAIN
    coo = pkg.CooMatrix.from_host(ctx, 4, 4, np.array([0, 1, 2, 4], np.int32),
                                  np.array([0, 1, 2, 3], np.int32), np.ones(4))
    with pytest.raises(pkg.B200Error):
        pkg.CsrMatrix(coo)


def test_bad_arguments_fail_loudly(ctx):
    L = pkg.lib()
    x = ctx.zeros(8, np.float64)
    assert L.b200_spmv_sell_f64(ctx.h, x.ptr, x.ptr, x.ptr, x.ptr, x.ptr, 16, 1, 8, None, None) == pkg.ERR_UNSUPPORTED
    assert L.b200_spmv_cmrs_f64(ctx.h, x.ptr, x.ptr, x.ptr, x.ptr, x.ptr, x.ptr, 1, 64, 8, None) == pkg.ERR_UNSUPPORTED
    assert L.b200_spmv_ellcm_f64(ctx.h, x.ptr, x.ptr, x.ptr, x.ptr, 8, 1, 8) == pkg.ERR_INVALID_VALUE
    assert L.b200_spmv_csr_f64(None, x.ptr, x.ptr, x.ptr, x.ptr, x.ptr, 1, None) == pkg.ERR_INVALID_VALUE
    assert b"null context" in L.b200_last_error()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_unaligned_arrays_take_the_scalar_kernels(ctx, dtype):
    """Arrays that are not 16-byte aligned (e.g. a shard view into a larger allocation) must still
    give the right answer through the scalar-load kernel variants."""
    n_rows, n_cols = 777, 900
    rows, cols, vals = random_sorted_matrix(n_rows, n_cols, 1, 50, 21)
    x = np.arange(n_cols, dtype=np.float64)
    y_ref = O.yref(n_rows, rows, cols, vals, x)
    nnz = rows.size
    L, suf, V = pkg.lib(), pkg.suffix(dtype), np.dtype(dtype).itemsize
    pad = lambda a, dt: np.concatenate([np.zeros(1, dt), a.astype(dt)])   # shift by one element
    rows_d, cols_d, vals_d = ctx.array(pad(rows, np.int32)), ctx.array(pad(cols, np.int32)), ctx.array(pad(vals, dtype))
    ptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n_rows))]).astype(np.int32)
    ptr_d, xd = ctx.array(ptr), ctx.array(x.astype(dtype))
    yd = ctx.array(np.full(n_rows, np.nan, dtype))
    pkg.check(getattr(L, "b200_spmv_csr_" + suf)(ctx.h, ptr_d.ptr, cols_d.ptr + 4, vals_d.ptr + V, xd.ptr,
                                                 yd.ptr, n_rows, None), "csr")
    check_y("csr-unaligned", yd.download(), y_ref, dtype)
    yd = ctx.array(np.full(n_rows, np.nan, dtype))
    pkg.check(getattr(L, "b200_spmv_coo_" + suf)(ctx.h, rows_d.ptr + 4, cols_d.ptr + 4, vals_d.ptr + V,
                                                 xd.ptr, yd.ptr, nnz, n_rows), "coo")
    check_y("coo-unaligned", yd.download(), y_ref, dtype)
    sp, ris = O.build_cmrs(n_rows, rows)
    sp_d, ris_d = ctx.array(sp), ctx.array(pad(ris, np.int32))
    yd = ctx.array(np.full(n_rows, np.nan, dtype))
    pkg.check(getattr(L, "b200_spmv_cmrs_" + suf)(ctx.h, vals_d.ptr + V, cols_d.ptr + 4, sp_d.ptr, ris_d.ptr + 4,
                                                  xd.ptr, yd.ptr, len(sp) - 1, 8, n_rows, None), "cmrs")
    check_y("cmrs-unaligned", yd.download(), y_ref, dtype)
    ri, sc, sd = O.build_sell(n_rows, rows, cols, vals)
    sc_d, sd_d, ri_d = ctx.array(pad(sc, np.int32)), ctx.array(pad(sd, dtype)), ctx.array(ri)
    yd = ctx.array(np.full(n_rows, np.nan, dtype))
    pkg.check(getattr(L, "b200_spmv_sell_" + suf)(ctx.h, sd_d.ptr + V, sc_d.ptr + 4, xd.ptr, yd.ptr, ri_d.ptr,
                                                  32, len(ri) - 1, n_rows, None, None), "sell")
    check_y("sell-unaligned", yd.download(), y_ref, dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_csr_stream_kernel(ctx, dtype):
    """The nnz-split CSR kernel (short rows): chosen automatically for mean <= 16 / max <= 256, and
    forced here on rows that span several 1024-entry tiles, with empty rows sitting exactly on tile
    boundaries and at the end of the matrix."""
    rng = np.random.default_rng(61)
    x = rng.uniform(-1, 1, 3000)
    # (a) automatic: 7-point-stencil-like rows
    rows, cols, vals = random_sorted_matrix(20000, 3000, 1, 12, 62)
    coo = pkg.CooMatrix.from_host(ctx, 20000, 3000, rows, cols, vals)
    csr = pkg.CsrMatrix(coo)
    assert csr.plan_info().stream_tiles == -(-rows.size // 1024)
    yd = ctx.array(np.full(20000, np.nan, dtype))
    csr.spmv(ctx.array(x.astype(dtype)), yd)
    check_y("csr-stream", yd.download(), O.yref(20000, rows, cols, vals, x), dtype)
    # (b) forced: lengths chosen so that rows end exactly on tile boundaries, followed by empty rows,
    # plus rows of 1000 entries that straddle two or three tiles
    lens = np.array([1024, 0, 0, 500, 524, 0, 1000, 1000, 1000, 48, 0, 3, 1, 1020, 7, 0, 0])
    assert np.cumsum(lens)[0] % 1024 == 0 and np.cumsum(lens)[4] % 1024 == 0
    n_rows, n_cols = lens.size, 3000
    rows = np.repeat(np.arange(n_rows, dtype=np.int32), lens)
    cols = np.concatenate([np.sort(rng.choice(n_cols, int(k), replace=False)) for k in lens]).astype(np.int32)
    vals = rng.uniform(-1, 1, rows.size)
    ctx.set_option("B200_CSR_STREAM", "1")
    coo = pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals)
    csr = pkg.CsrMatrix(coo)
    assert csr.plan_info().stream_tiles == -(-rows.size // 1024)
    yd = ctx.array(np.full(n_rows, np.nan, dtype))
    csr.spmv(ctx.array(x.astype(dtype)), yd)
    check_y("csr-stream-forced", yd.download(), O.yref(n_rows, rows, cols, vals, x), dtype)
    for g in (2, 4):        # G groups per thread: tiles of 2048 / 4096 entries (the power-law default is 2)
        ctx.set_option("B200_CSR_STREAM_G", g)
        csr_g = pkg.CsrMatrix(coo)
        info = csr_g.plan_info()
        assert info.stream_tile_entries == 1024 * g and info.stream_tiles == -(-rows.size // (1024 * g))
        yg = ctx.array(np.full(n_rows, np.nan, dtype))
        csr_g.spmv(ctx.array(x.astype(dtype)), yg)
        check_y(f"csr-stream G={g}", yg.download(), O.yref(n_rows, rows, cols, vals, x), dtype)
    ctx.set_option("B200_CSR_STREAM_G", None)
    ctx.set_option("B200_CSR_STREAM", "0")
    csr2 = pkg.CsrMatrix(coo)
    assert csr2.plan_info().stream_tiles == 0
    y2 = ctx.array(np.full(n_rows, np.nan, dtype))
    csr2.spmv(ctx.array(x.astype(dtype)), y2)
    check_y("csr-vector", y2.download(), O.yref(n_rows, rows, cols, vals, x), dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_cmrs_long_strips(ctx, dtype):
    """Strips longer than 8192 entries (a hub row of a power-law matrix) are split over several
    warps by the CMRS plan; with and without the plan the answer is the oracle's."""
    n_rows, n_cols = 600, 70000
    rows, cols, vals = random_sorted_matrix(n_rows, n_cols, 1, 10, 71,
                                            long_rows=((3, 40000), (4, 9000), (333, 20000), (599, 8193)))
    x = np.random.default_rng(4).uniform(-1, 1, n_cols)
    y_ref = O.yref(n_rows, rows, cols, vals, x)
    coo = pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals)
    ctx.set_option("B200_CMRS_STREAM", 0)      # this test is about the strip-splitting plan
    m = pkg.CmrsMatrix(pkg.CsrMatrix(coo))
    # strip 0 (~49 000 entries) -> 5 extra segments, strip 41 (~20 000) -> 2, strip 74 (~8 230) -> 1
    assert m.plan_extra_items() == 8 and m.plan_stream_tiles() == 0
    xd = ctx.array(x.astype(dtype))
    for use_plan in (True, False):
        yd = ctx.array(np.full(n_rows, np.nan, dtype))
        m.spmv(xd, yd, use_plan=use_plan)
        check_y(f"cmrs-long-plan{use_plan}", yd.download(), y_ref, dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("height", [8, 5, 32])
def test_cmrs_stream_kernel(ctx, dtype, height):
    """The nnz-split CMRS kernel (skewed strip lengths: warps own 512-entry tiles, not strips): picked
    by the plan on a hub-row matrix, forced on a ragged one; strips that span many tiles, tiles that
    span many strips, EMPTY strips (runs of empty rows) in the middle and at the end, both layouts
    (two arrays / packed) and both load-batch depths."""
    rng = np.random.default_rng(77)
    n_cols = 70000
    x = rng.uniform(-1, 1, n_cols)
    xd = ctx.array(x.astype(dtype))
    # (a) hub rows: chosen automatically
    n_rows = 600
    rows, cols, vals = random_sorted_matrix(n_rows, n_cols, 1, 10, 71,
                                            long_rows=((3, 40000), (4, 9000), (333, 20000), (599, 8193)))
    if height == 32:       # 19 strips of 32 rows: the hubs no longer stand out against the mean strip
        ctx.set_option("B200_CMRS_STREAM", 1)
    cmrs = pkg.CmrsMatrix(pkg.CsrMatrix(pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals)), height=height)
    assert cmrs.plan_stream_tiles() == -(-rows.size // 512)
    y_ref = O.yref(n_rows, rows, cols, vals, x)
    for u in (1, 2):
        ctx.set_option("B200_CMRS_U", u)
        for name, m in (("two-array", cmrs), ("packed", cmrs.packed())):
            yd = ctx.array(np.full(n_rows, np.nan, dtype))
            m.spmv(xd, yd)
            check_y(f"cmrs-stream {name} U={u}", yd.download(), y_ref, dtype)
    # (b) forced, with empty rows: 100 empty rows (= whole empty strips) after row 40, 3 short rows,
    # 500 empty rows at the end
    lens = np.concatenate([rng.integers(1, 90, 41), np.zeros(100, int), [700, 1, 2], rng.integers(0, 3, 300),
                           [3000], np.zeros(500, int)])
    n_rows = lens.size
    rows = np.repeat(np.arange(n_rows, dtype=np.int32), lens)
    cols = np.concatenate([np.sort(rng.choice(n_cols, int(k), replace=False)) for k in lens if k]).astype(np.int32)
    vals = rng.uniform(-1, 1, rows.size)
    ctx.set_option("B200_CMRS_STREAM", 1)
    cmrs = pkg.CmrsMatrix(pkg.CsrMatrix(pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals)), height=height)
    assert cmrs.plan_stream_tiles() == -(-rows.size // 512)
    yd = ctx.array(np.full(n_rows, np.nan, dtype))
    cmrs.spmv(xd, yd)
    check_y("cmrs-stream empty strips", yd.download(), O.yref(n_rows, rows, cols, vals, x), dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sell_narrow_chunks(ctx, dtype):
    """Chunks of at most 8 columns (stencil matrices) take the lane = row path of the SELL kernel: rows
    of 1..8 entries, a chunk of exactly 8 next to one of 9 (the wide path), a partly empty last chunk,
    sigma-sorted layouts, and the launch-overlap variant."""
    n_rows, n_cols = 5003, 9000
    rows, cols, vals = random_sorted_matrix(n_rows, n_cols, 1, 8, 55, long_rows=((100, 9), (4000, 40)))
    x = np.random.default_rng(56).uniform(-1, 1, n_cols)
    y_ref = O.yref(n_rows, rows, cols, vals, x)
    csr = pkg.CsrMatrix(pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals))
    xd = ctx.array(x.astype(dtype))
    for sigma in (1, 64, 4096):
        sell = pkg.SellMatrix(csr, dtype, sigma=sigma)
        for wpc in (None, 1):
            for overlap in (False, True):
                ctx.set_option("B200_SELL_WPC", wpc)
                ctx.set_launch_overlap(overlap)
                yd = ctx.array(np.full(n_rows, np.nan, dtype))
                sell.spmv(xd, yd)
                ctx.set_launch_overlap(True)       # the default of a library-owned queue
                check_y(f"sell-narrow sigma={sigma} wpc={wpc} overlap={overlap}", yd.download(), y_ref, dtype)


def test_used_column_blocks(ctx):
    """b200_used_column_blocks: which 2^k-column blocks of x a row block reads (what e2e callers upload)."""
    rng = np.random.default_rng(5)
    n_cols = 1 << 20
    cols = np.concatenate([rng.integers(5000, 9000, 1000), rng.integers(n_cols - 3000, n_cols, 500), [0]]).astype(np.int32)
    cd = ctx.array(cols)
    for k in (0, 8, 12, 20):
        used = np.full(((n_cols - 1) >> k) + 1, 7, np.uint8)
        pkg.check(pkg.lib().b200_used_column_blocks(ctx.h, cd.ptr, cols.size, n_cols, k, used.ctypes.data), "used blocks")
        want = np.zeros_like(used)
        want[np.unique(cols >> k)] = 1
        np.testing.assert_array_equal(used, want)


@pytest.mark.parametrize("height", [1, 2, 5, 8, 16, 32])
def test_cmrs_heights(ctx, height):
    n_rows, n_cols = 1234, 1500
    rows, cols, vals = random_sorted_matrix(n_rows, n_cols, 1, 30, 31)
    x = np.arange(n_cols, dtype=np.float64)
    y_ref = O.yref(n_rows, rows, cols, vals, x)
    coo = pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals)
    m = pkg.CmrsMatrix(pkg.CsrMatrix(coo), height=height)
    sp, ris = O.build_cmrs(n_rows, rows, height=height)
    np.testing.assert_array_equal(m.strip_ptr.download(), sp)
    np.testing.assert_array_equal(m.row_in_strip.download(), ris)
    for dtype in (np.float64, np.float32):
        yd = ctx.array(np.full(n_rows, np.nan, dtype))
        m.spmv(ctx.array(x.astype(dtype)), yd)
        check_y(f"cmrs-h{height}", yd.download(), y_ref, dtype)


@pytest.mark.parametrize("sigma", [1, 2, 32, 64, 100, 256, 1000, 4096, 5000, 1 << 20])
def test_sell_sigma(ctx, sigma):
    """SELL-C-sigma is new (the reference has no sigma): the GPU window sort must equal the oracle's
    specification -- stable, descending length inside each sigma window, perm[new] = old -- and the
    layout must reduce to the reference's at sigma = 1."""
    n_rows, n_cols = 5000, 5000
    rows, cols, vals = random_sorted_matrix(n_rows, n_cols, 1, 60, 41, long_rows=((77, 900), (4000, 400)))
    x = np.random.default_rng(5).uniform(-1, 1, n_cols)
    y_ref = O.yref(n_rows, rows, cols, vals, x)
    perm, sp, sc, sd = O.build_sell_sigma(n_rows, rows, cols, vals, sigma)
    coo = pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals)
    csr = pkg.CsrMatrix(coo)
    for dtype in (np.float64, np.float32):
        for wide in (False, True):
            m = pkg.SellMatrix(csr, dtype, sigma=sigma, wide=wide)
            if sigma > 1:
                np.testing.assert_array_equal(m.perm.download(), perm)
            np.testing.assert_array_equal(m.slice_ptr.download(), sp)
            np.testing.assert_array_equal(m.cols.download(), sc)
            assert m.data.download().tobytes() == sd.astype(dtype).tobytes()
            yd = ctx.array(np.full(n_rows, np.nan, dtype))
            m.spmv(ctx.array(x.astype(dtype)), yd)
            check_y(f"sell-sigma{sigma}", yd.download(), y_ref, dtype)
    if sigma == 1:
        ri, _, _ = O.build_sell(n_rows, rows, cols, vals)
        np.testing.assert_array_equal(sp, ri.astype(np.int64))
    else:
        # sorting can only shrink the padded size
        assert sp[-1] <= O.build_sell_sigma(n_rows, rows, cols, vals, 1)[1][-1]


def test_full_cant_shape(ctx, cant_dir):
    """BASELINE config 2: all five formats on the 62 451-row cant-shaped stand-in, x = ramp; COO from
    the column-major file (coo.c:43), the others from the row-sorted one (csr.c:43)."""
    n_rows, n_cols, rows, cols, vals = O.read_mtx(cant_dir / "databases" / "cant-sorted.mtx")
    n2, c2, rows_c, cols_c, vals_c = O.read_mtx(cant_dir / "databases" / "cant.mtx")
    assert (n_rows, rows.size) == (62451, 4325625)
    for dtype in (np.float64, np.float32):
        coo_s = pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals)
        order = np.lexsort((rows, cols))
        assert np.array_equal(rows[order], rows_c) and np.array_equal(cols[order], cols_c)
        m, y_ref = run_all_formats(ctx, n_rows, n_cols, rows, cols, vals, dtype, coo_order=order)
        info = m["csr"].plan_info()
        # mean row length 69.3 -> 8 lanes per row (~ 4 vector iterations per lane), no long rows
        assert info.lanes_per_row == 8 and info.n_long_rows == 0 and info.max_len == 81


def test_row_sharded_equals_unsharded(ctx):
    """G row blocks run one after the other on one GPU reproduce the unsharded run row for
    row (bit-identical for the alignment-independent kernels), cut points aligned to lcm(32, 8)."""
    n_rows, n_cols = 20000, 20000
    rows, cols, vals = random_sorted_matrix(n_rows, n_cols, 1, 80, 51)
    x = np.random.default_rng(9).uniform(-1, 1, n_cols)
    lens = np.bincount(rows, minlength=n_rows)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    for dtype in (np.float64, np.float32):
        xd = ctx.array(x.astype(dtype))
        full = pkg.build_all(pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals), dtype)
        y_full = {}
        for name, mat in full.items():
            yd = ctx.zeros(n_rows, dtype)
            mat.spmv(xd, yd)
            y_full[name] = yd.download()
        for G in (2, 4, 8):
            cuts = pkg.partition_rows(ptr, G, align=32)
            assert cuts[0] == 0 and cuts[-1] == n_rows and np.all(np.diff(cuts) > 0) and np.all(cuts[:-1] % 32 == 0)
            nnz_per = np.diff(ptr[cuts])
            assert nnz_per.max() <= 1.10 * nnz_per.mean()
            parts = {k: [] for k in full}
            for g in range(G):
                r0, r1 = cuts[g], cuts[g + 1]
                sel = slice(ptr[r0], ptr[r1])
                shard = pkg.build_all(pkg.CooMatrix.from_host(ctx, r1 - r0, n_cols, rows[sel] - r0, cols[sel], vals[sel]), dtype,
                                      ell=False)
                # same global width for ELL so that rows see identical padding
                shard["ell"] = pkg.EllMatrix(shard["csr"], dtype, row_size=full["ell"].row_size)
                shard["ellcm"] = pkg.EllCmMatrix(shard["csr"], dtype, row_size=full["ell"].row_size)
                for name, mat in shard.items():
                    yd = ctx.zeros(r1 - r0, dtype)
                    mat.spmv(xd, yd)
                    parts[name].append(yd.download())
            for name in full:
                got = np.concatenate(parts[name])
                if name in ("ell", "ellcm", "sell"):
                    # row-local AND alignment-independent (r0 % 32 == 0): identical bits
                    assert got.tobytes() == y_full[name].tobytes(), (name, G)
                else:
                    # csr/cmrs group entries by 16-byte alignment of the (rebased) entry index and
                    # coo uses atomics: same values up to summation order
                    assert O.rel_maxnorm(got, y_full[name].astype(np.float64)) <= TOL[np.dtype(dtype)]


VARIANT_HOOKS = {
    "csr": [{"B200_CSR_LANES": l, "B200_CSR_UNROLL": u} for l in (2, 4, 8, 16, 32) for u in (1, 2, 4)],
    "ell": [{"B200_ELL_LANES": l, "B200_ELL_UNROLL": u} for l in (2, 4, 8, 16, 32) for u in (1, 2, 4)],
    "sell": [{"B200_SELL_WPC": w, "B200_SELL_UNROLL": u} for w in (1, 2, 4, 8) for u in (1, 2, 4)],
    "coo": [{"B200_COO_U": u} for u in (1, 2, 4)],
    "cmrs": [{"B200_CMRS_U": u, "B200_CMRS_WPS": w} for u in (1, 2) for w in (1, 2, 4)],
}


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_every_tuning_variant_matches_the_oracle(ctx, dtype):
    """Every kernel variant the tuning hooks can select (lanes per row, load-batch depth U, SELL
    warps per chunk) computes the same y: ragged rows (1..150 entries), a row count that is not a
    multiple of 32 or 8, and a last SELL chunk / CMRS strip that is partly empty."""
    n_rows, n_cols = 2333, 4000
    rows, cols, vals = random_sorted_matrix(n_rows, n_cols, 1, 150, 71, long_rows=((7, 900), (2332, 1)))
    x = np.random.default_rng(72).uniform(-1, 1, n_cols)
    y_ref = O.yref(n_rows, rows, cols, vals, x)
    coo = pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals)
    m = pkg.build_all(coo, dtype)
    xd = ctx.array(x.astype(dtype))
    for fmt, envs in VARIANT_HOOKS.items():
        for env in envs:
            for k, v in env.items():
                ctx.set_option(k, str(v))
            mat = pkg.CsrMatrix(coo) if fmt == "csr" else m[fmt]   # the CSR plan caches its lanes
            if fmt == "csr":
                assert mat.plan_info().lanes_per_row == env["B200_CSR_LANES"]
            yd = ctx.array(np.full(n_rows, np.nan, dtype))
            mat.spmv(xd, yd)
            check_y(f"{fmt} {env}", yd.download(), y_ref, dtype)
            for k in env:
                ctx.set_option(k, None)


def test_launch_graph_replays_recorded_spmvs(ctx):
    """b200_graph_*: launches recorded on a context replay in order with one call, give the same
    bits as the direct launches, and calls that synchronise are refused while recording."""
    n_rows, n_cols = 3000, 3000
    rows, cols, vals = random_sorted_matrix(n_rows, n_cols, 1, 60, 81)
    x = np.random.default_rng(82).uniform(-1, 1, n_cols)
    coo = pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals)
    m = pkg.build_all(coo, np.float64)
    xd = ctx.array(x)
    direct, ys = {}, {}
    for name, mat in m.items():   # direct launches first: this also creates every plan
        yd = ctx.zeros(n_rows, np.float64)
        mat.spmv(xd, yd)
        direct[name] = yd.download()
        ys[name] = ctx.array(np.full(n_rows, np.nan))
    with ctx.record_graph() as g:
        for name, mat in m.items():
            mat.spmv(xd, ys[name])
    ctx.sync()
    assert all(np.isnan(ys[name].download()).all() for name in m), "recording must not execute"
    for _ in range(2):
        g.launch()
    for name in m:
        got = ys[name].download()
        if name == "coo":   # atomics: same values up to summation order
            assert O.rel_maxnorm(got, direct[name]) <= 1e-12
        else:
            assert got.tobytes() == direct[name].tobytes(), name
    # a call that has to synchronise (no plan -> statistics pass) fails loudly while recording
    csr = pkg.CsrMatrix(coo)
    with pytest.raises(pkg.B200Error):
        with ctx.record_graph():
            csr.spmv(xd, ys["csr"], use_plan=False)
    yd = ctx.zeros(n_rows, np.float64)   # and the context is usable afterwards
    m["csr"].spmv(xd, yd)
    assert yd.download().tobytes() == direct["csr"].tobytes()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("height", [8, 5, 32])
def test_cmrs_packed_layout(ctx, dtype, height):
    """Packed CMRS: (row_in_strip << 27) | column must equal the reference arrays bit for bit when
    unpacked, and the packed kernel must give the same bits as the two-array kernel (same order of
    operations) -- including the long-strip plan and both load-batch depths."""
    n_rows, n_cols = 3001, 5000
    rows, cols, vals = random_sorted_matrix(n_rows, n_cols, 1, 90, 91, long_rows=((11, 4000), (12, 4900)))
    x = np.random.default_rng(92).uniform(-1, 1, n_cols)
    y_ref = O.yref(n_rows, rows, cols, vals, x)
    coo = pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals)
    cmrs = pkg.CmrsMatrix(pkg.CsrMatrix(coo), height=height)
    packed = cmrs.packed()
    w = packed.packed.download().view(np.uint32)
    np.testing.assert_array_equal((w >> 27).astype(np.int32), cmrs.row_in_strip.download())
    np.testing.assert_array_equal((w & ((1 << 27) - 1)).astype(np.int32), cols)
    xd = ctx.array(x.astype(dtype))
    for u in (1, 2):
        ctx.set_option("B200_CMRS_U", str(u))
        y0, y1 = ctx.array(np.full(n_rows, np.nan, dtype)), ctx.array(np.full(n_rows, np.nan, dtype))
        cmrs.spmv(xd, y0)
        packed.spmv(xd, y1)
        check_y("cmrs-packed", y1.download(), y_ref, dtype)
        if cmrs.plan_extra_items() == 0:
            assert y1.download().tobytes() == y0.download().tobytes()
    assert packed.nbytes(dtype) == cmrs.nbytes(dtype) - 4 * rows.size
    # more than 2^27 columns cannot be packed
    with pytest.raises(pkg.B200Error):
        pkg.check(pkg.lib().b200_cmrs_pack(ctx.h, cmrs.cols.ptr, cmrs.row_in_strip.ptr, rows.size,
                                           (1 << 27) + 1, height, packed.packed.ptr), "b200_cmrs_pack")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_launch_overlap_chain_is_bit_identical(ctx, dtype):
    """b200_ctx_set_launch_overlap: a chain of DEPENDENT SpMVs (x -> y -> x -> ..., each launch reads
    what the previous one wrote, no other kernel in between) gives the same bits with the launches
    overlapped (programmatic dependent launch: matrix streamed before the wait, x gathered after it)
    as in order -- directly, and replayed from a launch graph.  A stale or early read of x would
    change the result."""
    n = 6000
    rows, cols, vals = random_sorted_matrix(n, n, 1, 70, 95)
    import scipy.sparse as sp
    A = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
    v = np.ones(n)
    for _ in range(30):                                  # scale to spectral radius ~1: 60 chained
        v = A @ v                                        # products stay in fp32 range
        rho = np.linalg.norm(v)
        v /= rho
    vals = vals / rho
    x0 = np.random.default_rng(96).uniform(-1, 1, n).astype(dtype)
    coo = pkg.CooMatrix.from_host(ctx, n, n, rows, cols, vals)
    m = pkg.build_all(coo, dtype)
    m["cmrs_packed"] = m["cmrs"].packed()
    steps = 60

    def chain(mat, a, b):
        for k in range(steps):
            src, dst = (a, b) if k % 2 == 0 else (b, a)
            mat.spmv(src, dst)

    for name, mat in m.items():
        a, b = ctx.array(x0), ctx.zeros(n, dtype)
        ctx.set_launch_overlap(False)
        chain(mat, a, b)                                 # in order (also creates the plans)
        want = a.download()
        assert np.isfinite(want).all() and np.abs(want).max() > 0
        ctx.set_launch_overlap(True)
        try:
            a2, b2 = ctx.array(x0), ctx.zeros(n, dtype)
            chain(mat, a2, b2)
            got = a2.download()
            a3, b3 = ctx.array(x0), ctx.zeros(n, dtype)
            with ctx.record_graph() as g:
                chain(mat, a3, b3)
            g.launch()
            got_graph = a3.download()
        finally:
            ctx.set_launch_overlap(True)                 # the default of a library-owned queue
        if name == "coo":                                # atomics: same values up to summation order
            scale = np.abs(want).max()
            assert np.abs(got.astype(np.float64) - want).max() <= 100 * TOL[np.dtype(dtype)] * scale
            assert np.abs(got_graph.astype(np.float64) - want).max() <= 100 * TOL[np.dtype(dtype)] * scale
        else:
            assert got.tobytes() == want.tobytes(), name
            assert got_graph.tobytes() == want.tobytes(), name


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_launch_overlap_is_fenced_by_everything_else(ctx, dtype):
    """Launch overlap is on by default on a library-owned queue.  It must be invisible: an SpMV issued
    right after an UPLOAD into a matrix array, a device-side BUILD or a memset must see the new data
    although back-to-back SpMV launches may start early.  A multi-megabyte upload lasts far longer than a
    kernel needs to start, so a kernel that streamed its matrix arrays early would read the old values."""
    n_rows, n_cols = 40000, 40000
    rows, cols, vals = random_sorted_matrix(n_rows, n_cols, 20, 60, 123)
    x = np.random.default_rng(124).uniform(-1, 1, n_cols)
    coo = pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals)
    xd = ctx.array(x.astype(dtype))
    yd = ctx.zeros(n_rows, dtype)
    csr = pkg.CsrMatrix(coo)
    csr.plan()
    vd = coo.values(dtype)
    rng = np.random.default_rng(125)
    for trial in range(6):
        for _ in range(3):                    # a chain of overlapping launches is in flight ...
            csr.spmv(xd, yd)
        new_vals = rng.uniform(-1, 1, vals.size)
        vd.upload(new_vals.astype(dtype))     # ... when the value array is overwritten (asynchronous copy)
        csr.spmv(xd, yd)                      # must read the NEW values
        check_y(f"csr after upload {trial}", yd.download(), O.yref(n_rows, rows, cols, new_vals, x), dtype)
        # a device-side build between two launches: the SELL fill writes the arrays the next launch reads
        coo.vals64.upload(new_vals)
        sell = pkg.SellMatrix(csr, dtype)     # b200_build_sell_* kernels, no sync
        sell.spmv(xd, yd)
        check_y(f"sell after build {trial}", yd.download(), O.yref(n_rows, rows, cols, new_vals, x), dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sell16_column_deltas_round_trip_and_spmv(ctx, dtype):
    """SELL-32 with 16-bit column deltas: chunk_base + delta16 must reproduce the reference's 32-bit
    indices bit for bit (padding slots -> the chunk base), the kernel must give the same bits as the
    32-bit SELL kernel (same order of operations), for every load-batch depth; a chunk spanning more
    than 65536 columns is refused."""
    n_rows, n_cols = 4099, 300000
    rng = np.random.default_rng(31)
    lens = rng.integers(1, 70, n_rows)
    centre = np.linspace(40000, n_cols - 40000, n_rows).astype(np.int64)
    rows = np.repeat(np.arange(n_rows, dtype=np.int32), lens)
    cols = np.concatenate([np.sort(rng.choice(np.arange(c - 30000, c + 30000) if r >= 64 else np.arange(1, 30000),
                                              int(k), replace=False))
                           for r, (c, k) in enumerate(zip(centre, lens))]).astype(np.int32)
    first = np.flatnonzero(rows == 0)             # a REAL entry in column 0, with a non-zero value
    cols[first] = np.sort(np.concatenate(([0], rng.choice(np.arange(1, 60000), first.size - 1, replace=False))))
    r40 = np.flatnonzero(rows == 40)              # a band that wraps around the matrix edge (periodic stencil)
    cols[r40] = np.sort(np.concatenate(([n_cols - 7], rng.choice(np.arange(0, 30000), r40.size - 1, replace=False))))
    r5 = np.flatnonzero(rows == 5)
    cols[r5] = np.sort(np.concatenate(([n_cols - 3], rng.choice(np.arange(0, 30000), r5.size - 1, replace=False))))
    vals = rng.uniform(-1, 1, rows.size)
    x = rng.uniform(-1, 1, n_cols)
    csr = pkg.CsrMatrix(pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals))
    sell = pkg.SellMatrix(csr, dtype)
    s16 = pkg.Sell16Matrix(sell)
    base = s16.chunk_base.download()
    d16 = s16.delta16.download()
    idx, dat, ri = sell.cols.download(), sell.data.download(), sell.row_indices.download()
    real = (idx != 0) | (dat != 0)
    chunk_of = np.repeat(np.arange(sell.n_slices), np.diff(ri))
    np.testing.assert_array_equal((base[chunk_of][real] + d16[real].astype(np.int64)) % n_cols, idx[real])
    assert np.all(d16[~real] == 0)
    for s in range(2, sell.n_slices):             # (chunks 0 and 1 wrap, see below) base = smallest real column
        sl = slice(ri[s], ri[s + 1])
        r = real[sl]
        assert base[s] == (idx[sl][r].min() if r.any() else 0)
    assert base[1] > n_cols // 2                  # row 40 reaches back past column 0: the chunk's band wraps
    xd = ctx.array(x.astype(dtype))
    y32 = ctx.array(np.full(n_rows, np.nan, dtype))
    ctx.set_option("B200_SELL_WPC", 1)
    for u in (1, 2, 4):
        ctx.set_option("B200_SELL_UNROLL", u)
        sell.spmv(xd, y32)
        y16 = ctx.array(np.full(n_rows, np.nan, dtype))
        s16.spmv(xd, y16)
        check_y(f"sell16 U={u}", y16.download(), O.yref(n_rows, rows, cols, vals, x), dtype)
        assert y16.download().tobytes() == y32.download().tobytes()
    assert s16.nbytes() == sell.nbytes() - 2 * sell.total + 4 * sell.n_slices
    # too wide: a row touching both ends of a 300 000-column matrix
    rows2, cols2, vals2 = random_sorted_matrix(64, n_cols, 2, 2, 32)
    cols2[0], cols2[1] = 100000, 200000
    wide = pkg.SellMatrix(pkg.CsrMatrix(pkg.CooMatrix.from_host(ctx, 64, n_cols, rows2, cols2, vals2)), dtype)
    with pytest.raises(pkg.B200Error) as e:
        pkg.Sell16Matrix(wide)
    assert e.value.status == pkg.ERR_UNSUPPORTED


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_conversions_back_to_csr(ctx, dtype):
    """ELL / SELL / CMRS -> CSR on the device: the row pointer, columns and values of the original CSR
    come back bit for bit (every format converts to every other through CSR and the builders)."""
    import ctypes as C
    L = pkg.lib()
    suf = pkg.suffix(dtype)
    n_rows, n_cols = 3001, 5000
    rows, cols, vals = random_sorted_matrix(n_rows, n_cols, 1, 40, 41, long_rows=((7, 300), (3000, 2)))
    m = pkg.build_all(pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals), dtype)
    ptr_ref = m["csr"].ptr.download()
    vals_t = vals.astype(dtype)

    def convert(fn, *args):
        ptr = ctx.empty(n_rows + 1, np.int32)
        nnz = C.c_longlong(0)
        pkg.check(fn(ctx.h, *args, ptr.ptr, None, None, C.byref(nnz)), "to_csr (count)")
        assert nnz.value == rows.size
        c, v = ctx.empty(nnz.value, np.int32), ctx.empty(nnz.value, dtype)
        pkg.check(fn(ctx.h, *args, ptr.ptr, c.ptr, v.ptr, None), "to_csr (fill)")
        np.testing.assert_array_equal(ptr.download(), ptr_ref)
        np.testing.assert_array_equal(c.download(), cols)
        assert v.download().tobytes() == vals_t.tobytes()

    ell, sell, cmrs = m["ell"], m["sell"], m["cmrs"]
    convert(getattr(L, "b200_ell_to_csr_" + suf), ell.data.ptr, ell.cols.ptr, n_rows, ell.row_size)
    convert(getattr(L, "b200_sell_to_csr_" + suf), sell.data.ptr, sell.cols.ptr, sell.row_indices.ptr, n_rows)
    ptr = ctx.empty(n_rows + 1, np.int32)
    pkg.check(L.b200_cmrs_to_csr_ptr(ctx.h, cmrs.strip_ptr.ptr, cmrs.row_in_strip.ptr, cmrs.n_strips, cmrs.height, n_rows,
                                     ptr.ptr), "b200_cmrs_to_csr_ptr")
    np.testing.assert_array_equal(ptr.download(), ptr_ref)


def test_format_advice(ctx):
    """b200_format_advice: per-format bytes equal the byte model of formats.py, regular matrices get the
    format with the fewest bytes, power-law ones SELL-sigma or CSR, never ELL."""
    # regular, uniform rows: CSR / SELL / ELL within a few bytes -> SELL
    rows, cols, vals = random_sorted_matrix(4096, 5000, 24, 24, 51)
    csr = pkg.CsrMatrix(pkg.CooMatrix.from_host(ctx, 4096, 5000, rows, cols, vals))
    a = csr.advice(np.float32)
    m = pkg.build_all(csr.coo, np.float32)
    for i, f in enumerate(pkg.FORMAT_NAMES):
        assert a.bytes[i] == m[f].nbytes(np.float32), f
    assert a.bytes_sell16 == pkg.Sell16Matrix(m["sell"]).nbytes()
    assert pkg.FORMAT_NAMES[a.recommended] == "sell" and not a.skewed and a.sell_padding == 1.0
    # ragged rows (1..60): ELL and SELL pad, CSR has the fewest bytes
    rows, cols, vals = random_sorted_matrix(4096, 5000, 1, 60, 52)
    a = pkg.CsrMatrix(pkg.CooMatrix.from_host(ctx, 4096, 5000, rows, cols, vals)).advice(np.float64)
    assert pkg.FORMAT_NAMES[a.recommended] == "csr" and a.bytes[pkg.FORMAT_ELL] > a.bytes[pkg.FORMAT_CSR]
    # hub rows: skewed -> SELL sigma-sorted when sorting removes the padding
    rows, cols, vals = random_sorted_matrix(70000, 70000, 1, 6, 53, long_rows=((5, 30000), (40000, 20000)))
    a = pkg.CsrMatrix(pkg.CooMatrix.from_host(ctx, 70000, 70000, rows, cols, vals)).advice(np.float32)
    assert a.skewed and pkg.FORMAT_NAMES[a.recommended] in ("sell", "csr") and pkg.FORMAT_NAMES[a.recommended] != "ell"
    assert a.sell_padding > a.sell_padding_sigma65536 >= 1.0
    assert (a.recommended_sigma == 65536) == (pkg.FORMAT_NAMES[a.recommended] == "sell")
    assert a.reason.decode().startswith("skewed rows")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_no_kernel_writes_outside_its_output(ctx, dtype):
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds WRITES are caught with guard bands:
    every format's y sits inside a larger buffer of sentinels (32 elements before, the rest after, so y
    starts 16-byte-misaligned for fp32 -- the kernels must not assume more than element alignment of y);
    after the SpMV the sentinels must be intact.  Row counts that are not multiples of 32 / 8, a partly
    empty last SELL chunk and CMRS strip, every launch variant the tuning hooks select."""
    SENT = -12345.0
    for n_rows, lo, hi in ((2333, 1, 150), (1001, 1, 8), (32, 3, 3), (5, 1, 300)):
        n_cols = 4000
        rows, cols, vals = random_sorted_matrix(n_rows, n_cols, lo, hi, n_rows)
        x = np.random.default_rng(n_rows).uniform(-1, 1, n_cols)
        y_ref = O.yref(n_rows, rows, cols, vals, x)
        m = pkg.build_all(pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals), dtype)
        m["cmrs_packed"] = m["cmrs"].packed()
        m["sell_sigma"] = pkg.SellMatrix(m["csr"], dtype, sigma=64)
        m["sell16"] = pkg.Sell16Matrix(m["sell"])
        xd = ctx.array(x.astype(dtype))
        pad_before, total = 33, n_rows + 33 + 4096
        for name, mat in m.items():
            variants = [{}]
            if name == "sell":
                variants = ([{"B200_SELL_WPC": w} for w in (1, 2, 4, 8)] + [{"B200_SELL_TMA": 1}] +
                            [{"B200_SELL_PIPE": k} for k in (2, 3, 4)])
            if name == "sell_sigma":
                variants = [{}, {"B200_SELL_PIPE": 3}]
            if name == "cmrs":
                variants = [{"B200_CMRS_WPS": w} for w in (1, 2, 4)] + [{"B200_CMRS_STREAM": 1}]
            if name == "csr":
                variants = [{}, {"B200_CSR_STREAM": 1}, {"B200_CSR_STREAM": 1, "B200_CSR_STREAM_G": 4}]
            for env in variants:
                for k, v in env.items():
                    ctx.set_option(k, v)
                if name == "csr":
                    mat = pkg.CsrMatrix(m["csr"].coo)
                if name == "cmrs" and "B200_CMRS_STREAM" in env:
                    mat = pkg.CmrsMatrix(m["csr"])
                buf = ctx.array(np.full(total, SENT, dtype))
                y = pkg.DeviceArray.from_ptr(ctx, buf.ptr + pad_before * np.dtype(dtype).itemsize, n_rows, dtype)
                mat.spmv(xd, y)
                got = buf.download()
                for k in env:
                    ctx.set_option(k, None)
                assert np.all(got[:pad_before] == SENT) and np.all(got[pad_before + n_rows:] == SENT), (name, env, n_rows)
                check_y(f"guarded {name} {env}", got[pad_before:pad_before + n_rows], y_ref, dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sell_pipelined_kernel_on_a_stencil_matrix(ctx, dtype):
    """A 7-point Laplacian large enough (80^3 rows = 16 000 chunks, none wider than 8 columns) that
    b200_spmv_sell_* with a plan takes the persistent pipelined kernel by itself; every forced variant
    (B200_SELL_PIPE = 0 | 2 | 3 | 4) must return the same y bit for bit, equal to the oracle's."""
    from test_distributed_cpu import laplace7
    n, rows, cols, vals = laplace7(80, 80, 80)
    x = np.random.default_rng(3).uniform(-1, 1, n)
    y_ref = O.yref(n, rows, cols, vals.astype(dtype).astype(np.float64), x.astype(dtype).astype(np.float64))
    coo = pkg.CooMatrix.from_host(ctx, n, n, rows, cols, vals)
    sell = pkg.SellMatrix(pkg.CsrMatrix(coo), dtype, wide=True)       # wide=True: the SpMV call carries a plan
    xd = ctx.array(x.astype(dtype))
    outs = {}
    for pipe in (None, 0, 2, 3, 4):
        ctx.set_option("B200_SELL_PIPE", pipe)
        yd = ctx.array(np.full(n, np.nan, dtype))
        sell.spmv(xd, yd)
        outs[pipe] = yd.download()
    for pipe, got in outs.items():
        assert np.array_equal(got, outs[0]), pipe                      # same products, same summation order
    check_y("sell pipelined stencil", outs[None], y_ref, dtype)
