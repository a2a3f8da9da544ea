"""CPU-only: the C-ABI library loads, exports every symbol include/b200spmv.h declares, the Python
signature table matches the header, and (without a GPU) device entry points fail loudly instead of
falling back to anything."""
import ctypes as C
import re
import subprocess

import pytest

from __graft_entry__ import load_package

pkg = load_package()


def declared_functions():
    text = pkg.HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ("b200_get_device_count", "b200_ctx_create", "b200_malloc", "b200_memcpy_h2d_async",
                 "b200_memcpy_d2h", "b200_sync", "b200_event_elapsed_ms",
                 "b200_spmv_coo_f64", "b200_spmv_csr_f32", "b200_spmv_ell_f64", "b200_spmv_ellcm_f32",
                 "b200_spmv_sell_f64", "b200_spmv_sell64_f32", "b200_spmv_cmrs_f64",
                 "b200_build_csr_ptr", "b200_build_sell_ptr", "b200_build_cmrs", "b200_partition_rows"):
        assert must in names


def test_library_exports_every_declared_symbol():
    assert pkg.LIB_PATH.exists(), "run __graft_entry__.build() first"
    out = subprocess.run(["nm", "-D", "--defined-only", str(pkg.LIB_PATH)], capture_output=True,
                         text=True, check=True).stdout
    exported = set(re.findall(r" T (b200_[a-z0-9_]+)", out))
    missing = [n for n in declared_functions() if n not in exported]
    assert not missing, f"declared in include/b200spmv.h but not exported: {missing}"


def test_python_signature_table_matches_header():
    assert sorted(pkg.SIGNATURES) == declared_functions()
    L = pkg.lib()  # resolves every symbol
    assert L.b200_version() == 100


def test_library_carries_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", str(pkg.LIB_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    n = C.c_int(-1)
    rc = pkg.lib().b200_get_device_count(C.byref(n))
    assert rc == pkg.ERR_NO_DEVICE and n.value == 0
    with pytest.raises(pkg.B200Error) as e:
        pkg.Context(0)
    assert e.value.status == pkg.ERR_NO_DEVICE
    assert pkg.lib().b200_status_string(rc) == b"no CUDA device"


def test_host_only_entry_points():
    """Slice/strip counts, the nnz-balanced partitioner and the host generator twin need no GPU."""
    import numpy as np
    L = pkg.lib()
    assert L.b200_sell_num_slices(62451, 32) == 1952      # sigma_c.c:74-81
    assert L.b200_cmrs_num_strips(62451, 8) == 7807       # cmrs.c:72
    assert L.b200_sell_num_slices(64, 32) == 2 and L.b200_cmrs_num_strips(64, 8) == 8
    lens = np.r_[np.full(512, 10), np.full(512, 30)]
    ptr = np.r_[0, np.cumsum(lens)].astype(np.int32)
    cuts = pkg.partition_rows(ptr, 2, align=32)
    assert cuts[0] == 0 and cuts[2] == 1024 and cuts[1] % 32 == 0
    half = ptr[-1] / 2
    assert abs(ptr[cuts[1]] - half) <= 32 * 30
    # 4 parts of a uniform matrix are equal
    ptr = (np.arange(4097) * 7).astype(np.int32)
    assert list(pkg.partition_rows(ptr, 4, align=32)) == [0, 1024, 2048, 3072, 4096]
    # banded generator: 64 distinct in-range columns per row, row-sorted
    n, npr = 10000, 64
    rows = np.empty(100 * npr, np.int32); cols = np.empty_like(rows); vals = np.empty(rows.size)
    assert L.b200_gen_banded_coo_host(n, 4950, 100, npr, 2000, 42, rows.ctypes.data,
                                      cols.ctypes.data, vals.ctypes.data) == 0
    assert np.array_equal(rows, np.repeat(np.arange(4950, 5050), npr))
    assert cols.min() >= 0 and cols.max() < n
    for r in range(100):
        assert len(set(cols[r * npr:(r + 1) * npr])) == npr
    assert np.all(np.abs(vals) <= 1.0) and np.all(vals != 0.0)
