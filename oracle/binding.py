"""ctypes binding of the CPU oracle (oracle/liboracle.so) and of the compiled reference
(oracle/_ref).  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product never does.

Nothing here reads /root/reference at run time -- it only loads files already built under
oracle/ (the GPU box has those, not the reference sources).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
FORMATS = ("coo", "csr", "ell", "sigma_c", "cmrs")
# kernel file each reference driver opens at run time (csr.c:136 etc.); content is irrelevant
# to the fake runtime but the file must exist or the driver exits with FileError.
KERNEL_FILES = {"coo": "Coo.cl", "csr": "Csr.cl", "ell": "Ell.cl", "sigma_c": "Sigma_C.cl",
                "cmrs": "Cmrs.cl"}
INPUT_FILES = {"coo": "cant.mtx", "csr": "cant-sorted.mtx", "ell": "cant-sorted.mtx",
               "sigma_c": "cant-sorted.mtx", "cmrs": "cant-sorted.mtx"}
# upload order = the arrays that define each format build
UPLOADS = {
    "csr": (("ptr", np.int32), ("cols", np.int32), ("data", np.float64), ("vect", np.float64)),
    "coo": (("rows", np.int32), ("cols", np.int32), ("data", np.float64), ("vect", np.float64)),
    "ell": (("data", np.float64), ("cols", np.int32), ("vect", np.float64)),
    "sigma_c": (("data", np.float64), ("cols", np.int32), ("vect", np.float64),
                ("row_indices", np.int32)),
    "cmrs": (("data", np.float64), ("cols", np.int32), ("strip_ptr", np.int32),
             ("row_in_strip", np.int32), ("vect", np.float64)),
}

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile the checker (liboracle.so, and oracle/_ref when /root/reference is present)."""
    if force or not (HERE / "liboracle.so").exists() or \
            (Path("/root/reference/csr.c").exists() and not (REF_DIR / "bin" / "csr").exists()):
        subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(str(HERE / "liboracle.so"))
    intp = C.POINTER(C.c_int)
    L.orc_mtx_read_size.argtypes = [C.c_char_p, intp, intp, intp]
    L.orc_mtx_read_coo.argtypes = [C.c_char_p, C.c_int, _i32p, _i32p, _f64p]
    L.orc_build_csr.argtypes = [C.c_int, C.c_int, _i32p, _i32p]
    L.orc_ell_stats.argtypes = [C.c_int, C.c_int, _i32p, intp, intp, intp]
    L.orc_build_ell.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f64p, _i32p, _f64p]
    L.orc_sell_num_slices.argtypes = [C.c_int, C.c_int]
    L.orc_build_sell_ptr.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p]
    L.orc_build_sell_ptr.restype = C.c_long
    L.orc_build_sell_fill.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f64p, _i32p,
                                      _i32p, _f64p]
    L.orc_cmrs_num_strips.argtypes = [C.c_int, C.c_int]
    L.orc_build_cmrs.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p]
    L.orc_sell_sigma_perm.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p]
    L.orc_build_sell_sigma.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f64p,
                                       _i32p, _i64p, C.c_void_p, C.c_void_p, C.c_long]
    L.orc_build_sell_sigma.restype = C.c_long
    L.orc_ell_to_colmajor.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _f64p, _i32p, _f64p]
    L.orc_set_threads.argtypes = [C.c_int]
    L.orc_yref_coo_serial.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p]
    for suf, fp in (("f64", _f64p), ("f32", _f32p)):
        getattr(L, f"orc_spmv_coo_{suf}").argtypes = [C.c_int, C.c_int, _i32p, _i32p, fp, fp, fp]
        getattr(L, f"orc_spmv_csr_{suf}").argtypes = [C.c_int, _i32p, _i32p, fp, fp, fp]
        getattr(L, f"orc_spmv_ell_{suf}").argtypes = [C.c_int, C.c_int, _i32p, fp, fp, fp]
        getattr(L, f"orc_spmv_sell_{suf}").argtypes = [C.c_int, C.c_int, _i32p, _i32p, fp, fp, fp]
        getattr(L, f"orc_spmv_sell64_{suf}").argtypes = [C.c_int, C.c_int, C.c_int, _i64p,
                                                         C.c_void_p, _i32p, fp, fp, fp]
        getattr(L, f"orc_spmv_cmrs_{suf}").argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p,
                                                       _i32p, fp, fp, fp]
    L.orc_rel_maxnorm_f64.argtypes = [C.c_int, _f64p, _f64p]
    L.orc_rel_maxnorm_f64.restype = C.c_double
    L.orc_rel_maxnorm_f32.argtypes = [C.c_int, _f32p, _f64p]
    L.orc_rel_maxnorm_f32.restype = C.c_double
    L.orc_check_abs.argtypes = [C.c_int, _f64p, _f64p, C.c_double]
    _lib = L
    return L


# ------------------------------------------------------------------------------------------
# numpy-level wrappers
# ------------------------------------------------------------------------------------------
def read_mtx(path):
    """(n_rows, n_cols, rows, cols, vals) with 0-based int32 indices in FILE ORDER."""
    L = lib()
    r, c, n = C.c_int(), C.c_int(), C.c_int()
    rc = L.orc_mtx_read_size(str(path).encode(), C.byref(r), C.byref(c), C.byref(n))
    if rc:
        raise IOError(f"orc_mtx_read_size({path}) -> {rc}")
    rows = np.empty(n.value, np.int32)
    cols = np.empty(n.value, np.int32)
    vals = np.empty(n.value, np.float64)
    rc = L.orc_mtx_read_coo(str(path).encode(), n.value, rows, cols, vals)
    if rc:
        raise IOError(f"orc_mtx_read_coo({path}) -> {rc}")
    return r.value, c.value, rows, cols, vals


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def build_csr(n_rows, rows):
    rows = _c(rows, np.int32)
    ptr = np.empty(n_rows + 1, np.int32)
    changes = lib().orc_build_csr(n_rows, rows.size, rows, ptr)
    return ptr, changes


def ell_stats(n_rows, rows):
    rows = _c(rows, np.int32)
    hi, lo, tot = C.c_int(), C.c_int(), C.c_int()
    last = lib().orc_ell_stats(n_rows, rows.size, rows, C.byref(hi), C.byref(lo), C.byref(tot))
    return hi.value, lo.value, tot.value, last


def build_ell(n_rows, row_size, rows, cols, vals):
    rows, cols, vals = _c(rows, np.int32), _c(cols, np.int32), _c(vals, np.float64)
    ec = np.empty(n_rows * row_size, np.int32)
    ed = np.empty(n_rows * row_size, np.float64)
    rc = lib().orc_build_ell(n_rows, rows.size, row_size, rows, cols, vals, ec, ed)
    if rc:
        raise ValueError("input is off the reference's well-defined ELL domain")
    return ec, ed


def ell_to_colmajor(n_rows, row_size, pitch, ec, ed):
    cc = np.empty(pitch * row_size, np.int32)
    cd = np.empty(pitch * row_size, np.float64)
    lib().orc_ell_to_colmajor(n_rows, row_size, pitch, _c(ec, np.int32), _c(ed, np.float64), cc, cd)
    return cc, cd


def build_sell(n_rows, rows, cols, vals, chunk=32):
    rows, cols, vals = _c(rows, np.int32), _c(cols, np.int32), _c(vals, np.float64)
    L = lib()
    ns = L.orc_sell_num_slices(n_rows, chunk)
    ri = np.empty(ns + 1, np.int32)
    total = L.orc_build_sell_ptr(n_rows, rows.size, chunk, rows, ri)
    sc = np.empty(total, np.int32)
    sd = np.empty(total, np.float64)
    rc = L.orc_build_sell_fill(n_rows, rows.size, chunk, rows, cols, vals, ri, sc, sd)
    if rc:
        raise ValueError("input is off the reference's well-defined SELL domain")
    return ri, sc, sd


def build_sell_sigma(n_rows, rows, cols, vals, sigma, chunk=32):
    rows, cols, vals = _c(rows, np.int32), _c(cols, np.int32), _c(vals, np.float64)
    L = lib()
    ns = L.orc_sell_num_slices(n_rows, chunk)
    perm = np.empty(n_rows, np.int32)
    sp = np.empty(ns + 1, np.int64)
    total = L.orc_build_sell_sigma(n_rows, rows.size, chunk, sigma, rows, cols, vals, perm, sp,
                                   None, None, 0)
    sc = np.empty(total, np.int32)
    sd = np.empty(total, np.float64)
    got = L.orc_build_sell_sigma(n_rows, rows.size, chunk, sigma, rows, cols, vals, perm, sp,
                                 sc.ctypes.data, sd.ctypes.data, total)
    assert got == total
    return perm, sp, sc, sd


def build_cmrs(n_rows, rows, height=8):
    rows = _c(rows, np.int32)
    L = lib()
    nt = L.orc_cmrs_num_strips(n_rows, height)
    sp = np.empty(nt + 1, np.int32)
    ris = np.empty(rows.size, np.int32)
    L.orc_build_cmrs(n_rows, rows.size, height, rows, sp, ris)
    return sp, ris


def yref(n_rows, rows, cols, vals, x):
    y = np.empty(n_rows, np.float64)
    lib().orc_yref_coo_serial(n_rows, len(rows), _c(rows, np.int32), _c(cols, np.int32),
                              _c(vals, np.float64), _c(x, np.float64), y)
    return y


def _suf(dtype):
    return "f32" if np.dtype(dtype) == np.float32 else "f64"


def spmv_coo(n_rows, rows, cols, vals, x, dtype=np.float64):
    y = np.empty(n_rows, dtype)
    getattr(lib(), "orc_spmv_coo_" + _suf(dtype))(n_rows, len(rows), _c(rows, np.int32),
                                                  _c(cols, np.int32), _c(vals, dtype),
                                                  _c(x, dtype), y)
    return y


def spmv_csr(n_rows, ptr, cols, vals, x, dtype=np.float64):
    y = np.empty(n_rows, dtype)
    getattr(lib(), "orc_spmv_csr_" + _suf(dtype))(n_rows, _c(ptr, np.int32), _c(cols, np.int32),
                                                  _c(vals, dtype), _c(x, dtype), y)
    return y


def spmv_ell(n_rows, row_size, cols, vals, x, dtype=np.float64):
    y = np.empty(n_rows, dtype)
    getattr(lib(), "orc_spmv_ell_" + _suf(dtype))(n_rows, row_size, _c(cols, np.int32),
                                                  _c(vals, dtype), _c(x, dtype), y)
    return y


def spmv_sell(row_indices, cols, vals, x, chunk=32, dtype=np.float64):
    ns = len(row_indices) - 1
    y = np.empty(ns * chunk, dtype)
    getattr(lib(), "orc_spmv_sell_" + _suf(dtype))(ns, chunk, _c(row_indices, np.int32),
                                                   _c(cols, np.int32), _c(vals, dtype),
                                                   _c(x, dtype), y)
    return y


def spmv_sell64(n_rows, slice_ptr, perm, cols, vals, x, chunk=32, dtype=np.float64):
    ns = len(slice_ptr) - 1
    y = np.zeros(n_rows, dtype)
    p = None if perm is None else _c(perm, np.int32)
    getattr(lib(), "orc_spmv_sell64_" + _suf(dtype))(n_rows, ns, chunk, _c(slice_ptr, np.int64),
                                                     None if p is None else p.ctypes.data,
                                                     _c(cols, np.int32), _c(vals, dtype),
                                                     _c(x, dtype), y)
    return y


def spmv_cmrs(n_rows, strip_ptr, row_in_strip, cols, vals, x, height=8, dtype=np.float64):
    y = np.empty(n_rows, dtype)
    getattr(lib(), "orc_spmv_cmrs_" + _suf(dtype))(n_rows, len(strip_ptr) - 1, height,
                                                   _c(strip_ptr, np.int32),
                                                   _c(row_in_strip, np.int32), _c(cols, np.int32),
                                                   _c(vals, dtype), _c(x, dtype), y)
    return y


def rel_maxnorm(y, y_ref):
    y_ref = _c(y_ref, np.float64)
    if np.asarray(y).dtype == np.float32:
        return lib().orc_rel_maxnorm_f32(len(y_ref), _c(y, np.float32), y_ref)
    return lib().orc_rel_maxnorm_f64(len(y_ref), _c(y, np.float64), y_ref)


# ------------------------------------------------------------------------------------------
# the compiled, unmodified reference (oracle/_ref)
# ------------------------------------------------------------------------------------------
def ref_available() -> bool:
    return all((REF_DIR / "bin" / f).exists() for f in FORMATS)


def prepare_ref_workdir(workdir: Path) -> None:
    """The reference drivers take no arguments and use relative paths (csr.c:43,136)."""
    (workdir / "databases").mkdir(parents=True, exist_ok=True)
    (workdir / "kernels").mkdir(parents=True, exist_ok=True)
    for name in KERNEL_FILES.values():
        p = workdir / "kernels" / name
        if not p.exists():
            p.write_text("/* placeholder: the fake OpenCL runtime never compiles this */\n")


def run_ref_driver(fmt: str, workdir: Path, threads: int | None = None):
    """Run oracle/_ref/bin/<fmt> in `workdir` (needs databases/<input>.mtx) with upload
    recording on.  Returns (exit_code, stdout, arrays, scalars, launch)."""
    prepare_ref_workdir(workdir)
    dump = workdir / f"dump_{fmt}"
    dump.mkdir(exist_ok=True)
    for old in dump.glob("*"):
        old.unlink()
    env = dict(os.environ, FAKECL_DUMP_DIR=str(dump), OMP_WAIT_POLICY="passive")
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    p = subprocess.run([str(REF_DIR / "bin" / fmt)], cwd=workdir, env=env, capture_output=True,
                       text=True, timeout=1200)
    arrays, scalars, launch = {}, {}, None
    man = dump / "manifest.txt"
    if man.exists():
        for line in man.read_text().splitlines():
            t = line.split()
            if t[0] == "upload":
                name, dt = UPLOADS[fmt][int(t[1])]
                arrays[name] = np.fromfile(dump / f"upload_{t[1]}.bin", dtype=dt)
            elif t[0] == "arg" and t[4] != "-":
                scalars[int(t[2])] = int(t[4])
            elif t[0] == "launch":
                launch = (int(t[2]), int(t[3]))
    return p.returncode, p.stdout, arrays, scalars, launch


_ref_libs = {}


def ref_lib(fmt: str, opt: bool = False) -> C.CDLL:
    key = (fmt, opt)
    if key not in _ref_libs:
        _ref_libs[key] = C.CDLL(str(REF_DIR / f"libref_{fmt}{'_O3' if opt else ''}.so"))
    return _ref_libs[key]


def ref_compute_using_cpu(fmt: str, arrays: dict, n_rows: int, nnz: int, opt: bool = False,
                          **kw) -> np.ndarray:
    """Call the reference's own compute_using_cpu (csr.c:285, coo.c:280, ell.c:357, cmrs.c:319)
    on numpy arrays.  The result buffer is zeroed first (the reference never does, quirk q2).
    NOTE: the symbol prints its 'CPU calculations' block to the C stdout of this process."""
    L = ref_lib(fmt, opt)
    f = L.compute_using_cpu
    f.restype = None
    out = np.zeros(n_rows + 8, np.float64)
    outp = C.c_void_p(out.ctypes.data)
    d = lambda k, dt: _c(arrays[k], dt).ctypes.data_as(C.c_void_p)
    keep = {k: _c(v, v.dtype) for k, v in arrays.items()}
    arrays = keep
    if fmt == "csr":
        f(d("data", np.float64), d("vect", np.float64), d("ptr", np.int32), d("cols", np.int32),
          C.c_int(n_rows), C.c_int(nnz), C.byref(outp))
    elif fmt == "coo":
        f(d("data", np.float64), d("vect", np.float64), d("rows", np.int32), d("cols", np.int32),
          C.c_int(nnz), C.byref(outp))
    elif fmt == "ell":
        f(d("data", np.float64), d("vect", np.float64), d("cols", np.int32), C.c_int(n_rows),
          C.c_int(kw["row_size"]), C.c_int(nnz), C.byref(outp))
    elif fmt == "cmrs":
        f(d("data", np.float64), d("vect", np.float64), d("strip_ptr", np.int32),
          d("row_in_strip", np.int32), d("cols", np.int32), C.c_int(len(arrays["strip_ptr"])),
          C.c_int(nnz), C.c_int(kw.get("height", 8)), C.byref(outp))
    else:
        raise ValueError(f"the reference has no CPU path for {fmt}")
    return out[:n_rows]


def ref_check_result(fmt: str, mtx_path, vect: np.ndarray, result: np.ndarray) -> bool:
    """The reference's own checker (inc/helper_functions.h:184-236), abs tolerance 1e-6."""
    f = ref_lib(fmt).check_result
    f.restype = C.c_bool
    f.argtypes = [C.c_char_p, _f64p, _f64p]
    return bool(f(str(mtx_path).encode(), _c(vect, np.float64), _c(result, np.float64)))
