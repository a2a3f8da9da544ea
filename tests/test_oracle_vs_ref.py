"""Pins the CPU oracle (oracle/spmv_oracle.c) against the UNMODIFIED reference compiled under
oracle/_ref: every format array the reference drivers upload must equal the oracle's build bit for
bit, and the reference's own compute_using_cpu / check_result must agree with the oracle's SpMV.

CPU only.  Skipped when oracle/_ref is absent (it is built from /root/reference by
oracle/Makefile; the committed fixtures in tests/golden cover that case)."""
import shutil

import numpy as np
import pytest

from oracle import binding as O
from conftest import random_sorted_matrix, write_mtx

pytestmark = pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")


def _check_all_formats(workdir, expect_rows=None):
    n_rows, n_cols, rows, cols, vals = O.read_mtx(workdir / "databases" / "cant-sorted.mtx")
    nnz = len(rows)
    if expect_rows:
        assert n_rows == expect_rows
    x = np.arange(n_cols, dtype=np.float64)  # csr.c:95-99
    y_ref = O.yref(n_rows, rows, cols, vals, x)

    # ---- CSR: csr.c:72-91 ----
    code, out, a, s, launch = O.run_ref_driver("csr", workdir)
    assert code == 0 and "cpu result is ok" in out
    ptr, changes = O.build_csr(n_rows, rows)
    assert changes == n_rows - 1
    np.testing.assert_array_equal(a["ptr"], ptr)
    np.testing.assert_array_equal(a["cols"], cols)
    assert a["data"].tobytes() == vals.tobytes()
    np.testing.assert_array_equal(a["vect"], x)
    assert s[5] == n_rows and launch == (8192, 256)
    y = O.spmv_csr(n_rows, ptr, cols, vals, x)
    assert O.rel_maxnorm(y, y_ref) <= 1e-12
    y_cpu = O.ref_compute_using_cpu("csr", a, n_rows, nnz)
    assert O.rel_maxnorm(y_cpu, y_ref) <= 1e-12
    assert O.ref_check_result("csr", workdir / "databases" / "cant-sorted.mtx", x, y_ref)

    # ---- ELL: ell.c:68-164 ----
    code, out, a, s, launch = O.run_ref_driver("ell", workdir)
    assert code == 0 and "cpu result is ok" in out
    hi, lo, tot, last = O.ell_stats(n_rows, rows)
    assert s[4] == n_rows and s[5] == hi and launch == (4096, 16)
    assert f"average column length {tot / n_rows:f}, shortest col {lo}, longest col {hi}" in out
    assert last <= hi, "matrix is off the reference's well-defined ELL domain"
    ec, ed = O.build_ell(n_rows, hi, rows, cols, vals)
    np.testing.assert_array_equal(a["cols"], ec)
    # the reference never writes padding DATA (malloc, ell.c:119): compare the real slots
    # bit-exactly and require the recording's padding to be whatever malloc gave (usually 0)
    real = np.zeros(n_rows * hi, bool)
    lens = np.bincount(rows, minlength=n_rows)
    for r in range(n_rows):
        real[r * hi:r * hi + lens[r]] = True
    assert a["data"][real].tobytes() == ed[real].tobytes()
    assert np.all(ed[~real] == 0.0)
    y = O.spmv_ell(n_rows, hi, ec, ed, x)
    assert O.rel_maxnorm(y, y_ref) <= 1e-12

    # ---- SELL-C (C=32, no sigma): sigma_c.c:71-202 ----
    code, out, a, s, launch = O.run_ref_driver("sigma_c", workdir)
    assert code == 0
    ri, sc, sd = O.build_sell(n_rows, rows, cols, vals)
    np.testing.assert_array_equal(a["row_indices"], ri)
    np.testing.assert_array_equal(a["cols"], sc)
    assert a["data"].tobytes() == sd.tobytes()
    ns = len(ri) - 1
    assert s[5] == 32 and launch == (ns * 32, 32)
    y = O.spmv_sell(ri, sc, sd, x)
    assert O.rel_maxnorm(y[:n_rows], y_ref) <= 1e-12
    assert np.all(y[n_rows:] == 0.0)
    # closed form: widths = max(1, longest row in slice)
    w = np.maximum.reduceat(lens, np.arange(0, n_rows, 32))
    np.testing.assert_array_equal(ri, np.concatenate([[0], np.cumsum(32 * np.maximum(w, 1))]))
    # the new sigma builder must reduce to the reference layout at sigma = 1
    perm, sp, sc2, sd2 = O.build_sell_sigma(n_rows, rows, cols, vals, sigma=1)
    np.testing.assert_array_equal(perm, np.arange(n_rows))
    np.testing.assert_array_equal(sp, ri.astype(np.int64))
    np.testing.assert_array_equal(sc2, sc)
    assert sd2.tobytes() == sd.tobytes()

    # ---- CMRS (height 8): cmrs.c:72-117 ----
    code, out, a, s, launch = O.run_ref_driver("cmrs", workdir)
    assert code == 0 and "cpu result is ok" in out
    sp, ris = O.build_cmrs(n_rows, rows)
    np.testing.assert_array_equal(a["strip_ptr"], sp)
    np.testing.assert_array_equal(a["row_in_strip"], ris)
    np.testing.assert_array_equal(a["cols"], cols)
    assert a["data"].tobytes() == vals.tobytes()
    assert s[6] == len(sp) - 1 and s[7] == 8 and launch == (8192, 32)
    np.testing.assert_array_equal(sp, np.concatenate([ptr[0:n_rows:8], [nnz]]))
    np.testing.assert_array_equal(ris, rows % 8)
    y = O.spmv_cmrs(n_rows, sp, ris, cols, vals, x)
    assert O.rel_maxnorm(y, y_ref) <= 1e-12
    y_cpu = O.ref_compute_using_cpu("cmrs", a, n_rows, nnz)
    assert O.rel_maxnorm(y_cpu, y_ref) <= 1e-12

    # ---- COO: coo.c:75-84 (reads cant.mtx, column-major file order) ----
    if (workdir / "databases" / "cant.mtx").exists():
        n2, c2, rows_c, cols_c, vals_c = O.read_mtx(workdir / "databases" / "cant.mtx")
        code, out, a, s, launch = O.run_ref_driver("coo", workdir)
        assert code == 0 and "cpu result is ok" in out
        np.testing.assert_array_equal(a["rows"], rows_c)
        np.testing.assert_array_equal(a["cols"], cols_c)
        assert a["data"].tobytes() == vals_c.tobytes()
        assert s[5] == len(rows_c) and launch == (-(-len(rows_c) // 64) * 64, 64)
        O.lib().orc_set_threads(1)
        y = O.spmv_coo(n2, rows_c, cols_c, vals_c, x)
        assert O.rel_maxnorm(y, O.yref(n2, rows_c, cols_c, vals_c, x)) <= 1e-12
        # same matrix, other order: the two y_ref agree to rounding
        assert O.rel_maxnorm(O.yref(n2, rows_c, cols_c, vals_c, x), y_ref) <= 1e-12


def test_small_fem(fem_small_dir):
    _check_all_formats(fem_small_dir, expect_rows=5 * 4 * 7 * 3)


@pytest.mark.parametrize("n_rows,lo,hi,seed", [(1, 1, 5, 0), (31, 1, 9, 1), (32, 1, 9, 2),
                                               (33, 1, 9, 3), (64, 2, 40, 4), (257, 1, 70, 5),
                                               (1000, 1, 3, 6)])
def test_random_ragged(tmp_path, n_rows, lo, hi, seed):
    """Ragged rows around the slice (32) and strip (8) boundaries; last row never the longest
    (the reference's ELL width ignores it, quirk q3)."""
    n_cols = max(n_rows, hi + 3)
    rows, cols, vals = random_sorted_matrix(n_rows, n_cols, lo, hi, seed,
                                            long_rows=[(0, hi), (n_rows - 1, lo)] if n_rows > 1 else [])
    if n_rows == 1:
        pytest.skip("single-row input: ELL width would be 0 in the reference (K ignores the last row)")
    (tmp_path / "databases").mkdir()
    write_mtx(tmp_path / "databases" / "cant-sorted.mtx", n_rows, n_cols, rows, cols, vals)
    order = np.lexsort((rows, cols))
    write_mtx(tmp_path / "databases" / "cant.mtx", n_rows, n_cols, rows[order], cols[order], vals[order])
    _check_all_formats(tmp_path, expect_rows=n_rows)


def test_symmetric_banner_is_ignored(fem_small_dir, tmp_path):
    """helper_functions.h:143-156 reads the banner and only rejects complex: a `symmetric` file is
    multiplied as stored (lower triangle only)."""
    import subprocess
    from conftest import gen_mtx_tool
    (tmp_path / "databases").mkdir()
    subprocess.run([str(gen_mtx_tool()), "--grid", "5", "4", "7", "--tri", "lower", "--banner",
                    "symmetric", "--out", str(tmp_path / "databases" / "cant-sorted.mtx")], check=True)
    _check_all_formats(tmp_path)


def test_full_cant_shape(cant_dir):
    """The 62 451-row stand-in at full size: every reference driver, every array."""
    _check_all_formats(cant_dir, expect_rows=62451)
