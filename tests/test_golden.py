"""Committed golden vectors: arrays RECORDED from the unmodified reference drivers
(tests/golden/make_golden.py) on a 180-row cant-shaped matrix.  They pin the oracle on boxes that
have no oracle/_ref (CPU test) and pin the GPU builders directly against the reference's own output
(GPU test), independently of the oracle."""
from pathlib import Path

import numpy as np
import pytest

from __graft_entry__ import load_package
from oracle import binding as O

G = np.load(Path(__file__).parent / "golden" / "ref_recording_fem_4x3x5.npz")
N_ROWS, N_COLS = (int(v) for v in G["shape"])
ROWS, COLS, VALS = G["input_sorted_rows"], G["input_sorted_cols"], G["input_sorted_vals"]
X = np.arange(N_COLS, dtype=np.float64)


def test_golden_is_a_reference_recording():
    assert N_ROWS == 4 * 3 * 5 * 3 and "cpu result is ok" in str(G["csr_stdout"])
    assert dict(G["csr_scalars"].tolist()) == {5: N_ROWS} and G["csr_launch"].tolist() == [8192, 256]
    assert dict(G["sigma_c_scalars"].tolist()) == {5: 32} and dict(G["cmrs_scalars"].tolist())[7] == 8
    np.testing.assert_array_equal(G["csr_vect"], X)


def test_oracle_matches_golden():
    ptr, _ = O.build_csr(N_ROWS, ROWS)
    np.testing.assert_array_equal(ptr, G["csr_ptr"])
    np.testing.assert_array_equal(COLS, G["csr_cols"])
    assert VALS.tobytes() == G["csr_data"].tobytes()
    hi, lo, tot, last = O.ell_stats(N_ROWS, ROWS)
    assert dict(G["ell_scalars"].tolist()) == {4: N_ROWS, 5: hi}
    ec, ed = O.build_ell(N_ROWS, hi, ROWS, COLS, VALS)
    np.testing.assert_array_equal(ec, G["ell_cols"])
    real = ed != 0.0
    assert ed[real].tobytes() == G["ell_data"][real].tobytes()
    ri, sc, sd = O.build_sell(N_ROWS, ROWS, COLS, VALS)
    np.testing.assert_array_equal(ri, G["sigma_c_row_indices"])
    np.testing.assert_array_equal(sc, G["sigma_c_cols"])
    assert sd.tobytes() == G["sigma_c_data"].tobytes()
    sp, ris = O.build_cmrs(N_ROWS, ROWS)
    np.testing.assert_array_equal(sp, G["cmrs_strip_ptr"])
    np.testing.assert_array_equal(ris, G["cmrs_row_in_strip"])
    np.testing.assert_array_equal(G["coo_rows"], G["input_colmajor_rows"])
    np.testing.assert_array_equal(G["coo_cols"], G["input_colmajor_cols"])
    # the reference's own CPU output (compute_using_cpu) against the oracle's y
    y_ref = O.yref(N_ROWS, ROWS, COLS, VALS, X)
    assert O.rel_maxnorm(G["csr_y_cpu"], y_ref) <= 1e-12
    assert O.rel_maxnorm(G["cmrs_y_cpu"], y_ref) <= 1e-12
    assert O.rel_maxnorm(O.spmv_csr(N_ROWS, ptr, COLS, VALS, X), G["csr_y_cpu"]) <= 1e-12


@pytest.mark.gpu
def test_gpu_builders_and_kernels_match_golden():
    pkg = load_package()
    ctx = pkg.Context(0)
    coo = pkg.CooMatrix.from_host(ctx, N_ROWS, N_COLS, ROWS, COLS, VALS)
    m = pkg.build_all(coo, np.float64)
    np.testing.assert_array_equal(m["csr"].ptr.download(), G["csr_ptr"])
    np.testing.assert_array_equal(m["ell"].cols.download(), G["ell_cols"])
    np.testing.assert_array_equal(m["sell"].row_indices.download(), G["sigma_c_row_indices"])
    np.testing.assert_array_equal(m["sell"].cols.download(), G["sigma_c_cols"])
    assert m["sell"].data.download().tobytes() == G["sigma_c_data"].tobytes()
    np.testing.assert_array_equal(m["cmrs"].strip_ptr.download(), G["cmrs_strip_ptr"])
    np.testing.assert_array_equal(m["cmrs"].row_in_strip.download(), G["cmrs_row_in_strip"])
    xd = ctx.array(X)
    for name, mat in m.items():
        yd = ctx.array(np.full(N_ROWS, np.nan))
        mat.spmv(xd, yd)
        assert O.rel_maxnorm(yd.download(), G["csr_y_cpu"]) <= 1e-12, name
    # COO in the reference's own (column-major file) order
    coo_c = pkg.CooMatrix.from_host(ctx, N_ROWS, N_COLS, G["coo_rows"], G["coo_cols"], G["coo_data"])
    yd = ctx.array(np.full(N_ROWS, np.nan))
    coo_c.spmv(xd, yd)
    assert O.rel_maxnorm(yd.download(), G["csr_y_cpu"]) <= 1e-12
    ctx.close()
