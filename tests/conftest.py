"""pytest configuration: markers, repo-root imports, shared matrix fixtures."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
os.environ.setdefault("OMP_WAIT_POLICY", "passive")

PKG = ROOT / "opencl-spmv-algorithms_b200"
GEN = PKG / "tools" / "gen_mtx"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a device: skip them (loudly) when there is none instead of failing."""
    try:
        from __graft_entry__ import load_package
        import ctypes
        n = ctypes.c_int(0)
        have_gpu = load_package().lib().b200_get_device_count(ctypes.byref(n)) == 0 and n.value > 0
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (runs on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def gen_mtx_tool() -> Path:
    if not GEN.exists() or GEN.stat().st_mtime < (PKG / "tools" / "gen_mtx.c").stat().st_mtime:
        subprocess.run(["gcc", "-O2", "-o", str(GEN), str(PKG / "tools" / "gen_mtx.c")], check=True)
    return GEN


def write_mtx(path, n_rows, n_cols, rows, cols, vals, banner="general"):
    """MatrixMarket coordinate text, entries in the given order, 1-based, values as %.17g."""
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate real {banner}\n% test matrix\n")
        f.write(f"{n_rows} {n_cols} {len(rows)}\n")
        for r, c, v in zip(rows, cols, vals):
            f.write(f"{int(r) + 1} {int(c) + 1} {float(v):.17g}\n")


def random_sorted_matrix(n_rows, n_cols, min_len, max_len, seed, long_rows=()):
    """Row-sorted triples, no empty rows, first row 0: the reference's well-defined domain.
    `long_rows` = iterable of (row, length) overrides.  Columns ascending within a row."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(min_len, max_len + 1, size=n_rows)
    for r, ln in long_rows:
        lens[r] = ln
    lens = np.minimum(np.maximum(lens, 1), n_cols)
    rows = np.repeat(np.arange(n_rows, dtype=np.int32), lens)
    cols = np.concatenate([np.sort(rng.choice(n_cols, size=int(ln), replace=False)) for ln in lens])
    vals = np.round(rng.uniform(-1.0, 1.0, size=rows.size), 6)
    vals[vals == 0.0] = 1e-6
    return rows.astype(np.int32), cols.astype(np.int32), vals.astype(np.float64)


@pytest.fixture(scope="session")
def fem_small_dir(tmp_path_factory):
    """A 5x4x7 grid x 3 dof cant-shaped matrix (420 rows): both file orders, full + lower."""
    d = tmp_path_factory.mktemp("fem_small")
    (d / "databases").mkdir()
    g = ["--grid", "5", "4", "7", "--dof", "3"]
    subprocess.run([str(gen_mtx_tool()), *g, "--order", "row", "--out",
                    str(d / "databases" / "cant-sorted.mtx")], check=True)
    subprocess.run([str(gen_mtx_tool()), *g, "--order", "col", "--out",
                    str(d / "databases" / "cant.mtx")], check=True)
    return d


@pytest.fixture(scope="session")
def cant_dir(tmp_path_factory):
    """Full-size cant-shaped stand-in (62 451 rows, 4 325 625 nnz), both file orders."""
    d = tmp_path_factory.mktemp("cant")
    (d / "databases").mkdir()
    subprocess.run([str(gen_mtx_tool()), "--order", "row", "--out",
                    str(d / "databases" / "cant-sorted.mtx")], check=True)
    subprocess.run([str(gen_mtx_tool()), "--order", "col", "--out",
                    str(d / "databases" / "cant.mtx")], check=True)
    return d
