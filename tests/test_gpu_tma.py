"""The bulk-copy (TMA engine) staged SELL kernel, selected with B200_SELL_TMA=1: cp.async.bulk into a
per-warp two-stage shared-memory ring, mbarrier completion, persistent grid
(csrc/spmv_sell_ell.cu: sell32_tma_kernel).  Same parity bar as every other kernel."""
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


CASES = [  # n_rows, n_cols, min_len, max_len, seed, long_rows
    (1, 40, 3, 3, 0, ()),                       # one chunk, one partial piece
    (33, 64, 1, 9, 1, ()),                      # two chunks, second almost empty
    (2333, 4000, 1, 150, 2, ((7, 900),)),       # chunks of 1..57 pieces, ragged
    (40000, 50000, 10, 40, 3, ()),              # more chunks than resident warps: the persistent loop
    (5000, 9000, 16, 16, 4, ()),                # every chunk exactly one full piece (16 columns)
    (5000, 9000, 32, 32, 5, ()),                # exactly two full pieces: ring wraps on chunk ends
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", CASES, ids=[f"r{c[0]}-l{c[2]}-{c[3]}" for c in CASES])
def test_sell_tma_matches_oracle(ctx, case, dtype):
    n_rows, n_cols, lo, hi, seed, long_rows = case
    rows, cols, vals = random_sorted_matrix(n_rows, n_cols, lo, hi, seed, long_rows)
    x = np.random.default_rng(seed + 100).uniform(-1, 1, n_cols)
    y_ref = O.yref(n_rows, rows, cols, vals, x)
    coo = pkg.CooMatrix.from_host(ctx, n_rows, n_cols, rows, cols, vals)
    csr = pkg.CsrMatrix(coo)
    xd = ctx.array(x.astype(dtype))
    for sigma in (1, 64):
        sell = pkg.SellMatrix(csr, dtype, sigma=sigma)
        y_plain = ctx.array(np.full(n_rows, np.nan, dtype))
        sell.spmv(xd, y_plain)
        for blocks in (1, 2):
            ctx.set_option("B200_SELL_TMA", "1")
            ctx.set_option("B200_SELL_TMA_BLOCKS", str(blocks))
            yd = ctx.array(np.full(n_rows, np.nan, dtype))
            sell.spmv(xd, yd, use_plan=False)   # no wide-chunk plan: whole chunks go through the ring
            got = yd.download()      # b200_memcpy_d2h also reports a timed-out mbarrier wait
            ctx.set_option("B200_SELL_TMA", None)
            ctx.set_option("B200_SELL_TMA_BLOCKS", None)
            err = O.rel_maxnorm(got, y_ref)
            assert err <= TOL[np.dtype(dtype)], (sigma, blocks, err)
            assert O.rel_maxnorm(got, y_plain.download().astype(np.float64)) <= 2 * TOL[np.dtype(dtype)]


def test_sell_tma_in_a_launch_graph(ctx):
    rows, cols, vals = random_sorted_matrix(3000, 3000, 1, 60, 9)
    x = np.random.default_rng(10).uniform(-1, 1, 3000)
    coo = pkg.CooMatrix.from_host(ctx, 3000, 3000, rows, cols, vals)
    sell = pkg.SellMatrix(pkg.CsrMatrix(coo), np.float64)
    xd = ctx.array(x)
    ctx.set_option("B200_SELL_TMA", "1")
    y1 = ctx.zeros(3000, np.float64)
    sell.spmv(xd, y1)
    y2 = ctx.array(np.full(3000, np.nan))
    with ctx.record_graph() as g:
        sell.spmv(xd, y2)
    g.launch()
    assert y2.download().tobytes() == y1.download().tobytes()
