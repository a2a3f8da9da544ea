"""Regenerates tests/golden/ref_recording_fem_4x3x5.npz.

Runs the UNMODIFIED reference drivers (oracle/_ref/bin/<fmt>, built from /root/reference by
oracle/Makefile against the recording fake OpenCL runtime) on a small cant-shaped matrix produced
by tools/gen_mtx (grid 4x3x5, 3 dof: 180 rows) and stores, per driver, every array it uploads and
its scalar kernel arguments.  The result is DATA recorded from running the reference on our own
input -- no reference source is copied.  Only runs where /root/reference (hence oracle/_ref) exists:

    python tests/golden/make_golden.py
"""
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import binding as O  # noqa: E402

GRID = ("4", "3", "5")


def main():
    assert O.ref_available(), "build oracle/_ref first (make -C oracle)"
    gen = ROOT / "opencl-spmv-algorithms_b200" / "tools" / "gen_mtx"
    out = {}
    with tempfile.TemporaryDirectory() as d:
        d = Path(d)
        (d / "databases").mkdir()
        for order, name in (("row", "cant-sorted.mtx"), ("col", "cant.mtx")):
            subprocess.run([str(gen), "--grid", *GRID, "--dof", "3", "--order", order, "--out",
                            str(d / "databases" / name)], check=True)
        for which, name in (("sorted", "cant-sorted.mtx"), ("colmajor", "cant.mtx")):
            n_rows, n_cols, rows, cols, vals = O.read_mtx(d / "databases" / name)
            out[f"input_{which}_rows"], out[f"input_{which}_cols"], out[f"input_{which}_vals"] = rows, cols, vals
        out["shape"] = np.array([n_rows, n_cols], np.int32)
        for fmt in O.FORMATS:
            code, stdout, arrays, scalars, launch = O.run_ref_driver(fmt, d)
            assert code == 0, (fmt, stdout)
            for k, v in arrays.items():
                out[f"{fmt}_{k}"] = v
            out[f"{fmt}_scalars"] = np.array(sorted(scalars.items()), np.int64)
            out[f"{fmt}_launch"] = np.array(launch, np.int64)
            out[f"{fmt}_stdout"] = np.array(stdout)
        # the reference's own CPU results (compute_using_cpu) for the formats that have one
        x = np.arange(n_cols, dtype=np.float64)
        nnz = len(out["input_sorted_rows"])
        a = {"ptr": out["csr_ptr"], "cols": out["csr_cols"], "data": out["csr_data"], "vect": x}
        out["csr_y_cpu"] = O.ref_compute_using_cpu("csr", a, n_rows, nnz)
        a = {"strip_ptr": out["cmrs_strip_ptr"], "row_in_strip": out["cmrs_row_in_strip"],
             "cols": out["cmrs_cols"], "data": out["cmrs_data"], "vect": x}
        out["cmrs_y_cpu"] = O.ref_compute_using_cpu("cmrs", a, n_rows, nnz)
    path = Path(__file__).with_name("ref_recording_fem_4x3x5.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, path.stat().st_size, "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
