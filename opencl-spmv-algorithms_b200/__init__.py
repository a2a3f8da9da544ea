"""b200spmv -- Python host side over the C ABI of libb200spmv.so (include/b200spmv.h).

The product is the CUDA library and the C drivers in host/; this module is the thin ctypes
mirror used by tests/, bench.py and __graft_entry__.py.  It never computes anything itself and has
no CPU fallback: if the shared library is missing, or there is no CUDA device, calls raise.

The directory name contains '-', so import it through `__graft_entry__.load_package()` (which
registers it as module `spmv_b200`).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libb200spmv.so"
HEADER_PATH = PKG_DIR.parent / "include" / "b200spmv.h"

SUCCESS = 0
ERR_NO_DEVICE, ERR_CUDA, ERR_INVALID_VALUE, ERR_OOM, ERR_UNSUPPORTED, ERR_DOMAIN = 1, 2, 3, 4, 5, 6


class B200Error(RuntimeError):
    def __init__(self, status: int, where: str, detail: str):
        super().__init__(f"{where}: status {status}: {detail}")
        self.status = status


class CsrPlanInfo(C.Structure):
    _fields_ = [("n_rows", C.c_int), ("nnz", C.c_longlong), ("min_len", C.c_int),
                ("max_len", C.c_int), ("mean_len", C.c_double), ("lanes_per_row", C.c_int),
                ("long_threshold", C.c_int), ("n_long_rows", C.c_int), ("stream_tiles", C.c_int),
                ("stream_tile_entries", C.c_int)]


class RowStats(C.Structure):
    _fields_ = [("max_len", C.c_int), ("min_len", C.c_int), ("sum_len", C.c_longlong),
                ("max_len_excl_last", C.c_int), ("min_len_excl_last", C.c_int),
                ("sum_len_excl_last", C.c_longlong), ("last_len", C.c_int)]


class BlockF64(C.Structure):
    """b200_block_f64: one rank's row block for the iterated mode."""
    _fields_ = [("format", C.c_int), ("n_rows", C.c_int), ("n_slices", C.c_int), ("ptr", C.c_void_p),
                ("indices", C.c_void_p), ("data", C.c_void_p), ("csr_plan", C.c_void_p)]


class IterDesc(C.Structure):
    """b200_iter_desc."""
    _fields_ = [("mode", C.c_int), ("world", C.c_int), ("rank", C.c_int), ("rows_per_rank", C.c_longlong),
                ("x", C.POINTER(C.c_void_p) * 2), ("halo_lo", C.POINTER(C.c_int)), ("halo_hi", C.POINTER(C.c_int)),
                ("graph_steps", C.c_int), ("mcast", C.c_void_p)]


class FormatAdvice(C.Structure):
    """b200_format_advice_t."""
    _fields_ = [("n_rows", C.c_int), ("nnz", C.c_longlong), ("min_len", C.c_int), ("max_len", C.c_int),
                ("mean_len", C.c_double), ("skewed", C.c_int), ("sell_padding", C.c_double),
                ("sell_padding_sigma65536", C.c_double), ("bytes", C.c_longlong * 5),
                ("bytes_sell_sigma65536", C.c_longlong), ("bytes_sell16", C.c_longlong), ("recommended", C.c_int),
                ("recommended_sigma", C.c_int), ("reason", C.c_char * 160)]


FORMAT_COO, FORMAT_CSR, FORMAT_ELL, FORMAT_SELL, FORMAT_CMRS = range(5)
FORMAT_NAMES = ("coo", "csr", "ell", "sell", "cmrs")
ITER_FUSED, ITER_ALLGATHER, ITER_FUSED_MCAST = 0, 1, 2
COMM_ID_BYTES = 128

_vp, _i, _ll, _u64, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_uint64, C.c_size_t
_vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); the single source of truth for the ABI seen from Python.
# tests/test_abi.py cross-checks this table against include/b200spmv.h.
SIGNATURES = {
    "b200_status_string": (C.c_char_p, [_i]),
    "b200_last_error": (C.c_char_p, []),
    "b200_version": (_i, []),
    "b200_get_device_count": (_i, [C.POINTER(_i)]),
    "b200_device_name": (_i, [_i, C.c_char_p, _sz]),
    "b200_device_sm_count": (_i, [_i, C.POINTER(_i)]),
    "b200_ctx_create": (_i, [_i, _vpp]),
    "b200_ctx_create_on_stream": (_i, [_i, _vp, _vpp]),
    "b200_ctx_destroy": (_i, [_vp]),
    "b200_ctx_device": (_i, [_vp, C.POINTER(_i)]),
    "b200_ctx_set_l2_persist": (_i, [_vp, _vp, _sz]),
    "b200_ctx_set_option": (_i, [_vp, C.c_char_p, C.c_char_p]),
    "b200_malloc": (_i, [_vp, _sz, _vpp]),
    "b200_free": (_i, [_vp, _vp]),
    "b200_memcpy_h2d_async": (_i, [_vp, _vp, _vp, _sz]),
    "b200_memcpy_d2h": (_i, [_vp, _vp, _vp, _sz]),
    "b200_memcpy_d2h_async": (_i, [_vp, _vp, _vp, _sz]),
    "b200_memcpy_d2d_async": (_i, [_vp, _vp, _vp, _sz]),
    "b200_memset_async": (_i, [_vp, _vp, _i, _sz]),
    "b200_sync": (_i, [_vp]),
    "b200_host_alloc_pinned": (_i, [_sz, _vpp]),
    "b200_host_free_pinned": (_i, [_vp]),
    "b200_event_create": (_i, [_vp, _vpp]),
    "b200_event_record": (_i, [_vp, _vp]),
    "b200_event_elapsed_ms": (_i, [_vp, _vp, C.POINTER(C.c_float)]),
    "b200_event_destroy": (_i, [_vp]),
    "b200_ctx_set_launch_overlap": (_i, [_vp, _i]),
    "b200_graph_begin": (_i, [_vp]),
    "b200_graph_end": (_i, [_vp, _vpp]),
    "b200_graph_launch": (_i, [_vp, _vp]),
    "b200_graph_destroy": (_i, [_vp]),
    "b200_csr_plan_create": (_i, [_vp, _vp, _i, _vpp]),
    "b200_csr_plan_get_info": (_i, [_vp, C.POINTER(CsrPlanInfo)]),
    "b200_csr_plan_destroy": (_i, [_vp]),
    "b200_spmv_csr_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "b200_spmv_csr_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "b200_spmv_coo_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i]),
    "b200_spmv_coo_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i]),
    "b200_spmv_ell_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i]),
    "b200_spmv_ell_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i]),
    "b200_spmv_ellcm_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i]),
    "b200_spmv_ellcm_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i]),
    "b200_sell_plan_create": (_i, [_vp, _vp, _i, _vpp]),
    "b200_sell64_plan_create": (_i, [_vp, _vp, _i, _vpp]),
    "b200_sell_plan_extra_items": (_i, [_vp, C.POINTER(_i)]),
    "b200_sell_plan_destroy": (_i, [_vp]),
    "b200_spmv_sell_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "b200_spmv_sell_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "b200_spmv_sell64_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "b200_spmv_sell64_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "b200_cmrs_plan_create": (_i, [_vp, _vp, _i, _vpp]),
    "b200_cmrs_plan_extra_items": (_i, [_vp, C.POINTER(_i)]),
    "b200_cmrs_plan_stream_tiles": (_i, [_vp, C.POINTER(_i)]),
    "b200_cmrs_plan_destroy": (_i, [_vp]),
    "b200_spmv_cmrs_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "b200_spmv_cmrs_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "b200_cmrs_pack": (_i, [_vp, _vp, _vp, _ll, _i, _i, _vp]),
    "b200_spmv_cmrs_packed_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "b200_spmv_cmrs_packed_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "b200_check_sorted_rows": (_i, [_vp, _vp, _i, _i]),
    "b200_build_csr_ptr": (_i, [_vp, _vp, _i, _i, _vp]),
    "b200_row_length_stats": (_i, [_vp, _vp, _i, C.POINTER(RowStats)]),
    "b200_build_ell_f64": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "b200_build_ell_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "b200_build_ell_colmajor_f64": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "b200_build_ell_colmajor_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "b200_sell_num_slices": (_i, [_i, _i]),
    "b200_build_sell_ptr": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, C.POINTER(_ll)]),
    "b200_sell_ptr_to_i32": (_i, [_vp, _vp, _i, _vp]),
    "b200_build_sell_fill_f64": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "b200_build_sell_fill_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "b200_cmrs_num_strips": (_i, [_i, _i]),
    "b200_build_cmrs": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "b200_convert_f64_to_f32": (_i, [_vp, _vp, _vp, _ll]),
    "b200_offset_i32": (_i, [_vp, _vp, _ll, _i]),
    "b200_fill_ramp_f64": (_i, [_vp, _vp, _i]),
    "b200_fill_ramp_f32": (_i, [_vp, _vp, _i]),
    "b200_gen_banded_nnz": (_ll, [_ll, _i, _i, _i]),
    "b200_gen_banded_coo": (_i, [_vp, _i, _i, _i, _i, _i, _u64, _vp, _vp, _vp]),
    "b200_gen_laplace7_nnz": (_ll, [_i, _i, _i, _i, _i]),
    "b200_gen_laplace7_coo": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "b200_gen_rmat_count": (_i, [_vp, _i, _i, C.c_double, C.c_double, C.c_double, _u64, _i, _i,
                                 C.POINTER(_ll)]),
    "b200_gen_rmat_coo": (_i, [_vp, _i, _i, C.c_double, C.c_double, C.c_double, _u64, _i, _i, _ll,
                               _vp, _vp, _vp, C.POINTER(_ll)]),
    "b200_gen_uniform_f64": (_i, [_vp, _vp, _ll, _u64, C.c_double, C.c_double]),
    "b200_gen_uniform_f32": (_i, [_vp, _vp, _ll, _u64, C.c_float, C.c_float]),
    "b200_gen_banded_coo_host": (_i, [_i, _i, _i, _i, _i, _u64, _vp, _vp, _vp]),
    "b200_gen_uniform_f64_host": (_i, [_vp, _ll, _u64, C.c_double, C.c_double]),
    "b200_partition_rows": (_i, [_vp, _i, _i, _i, _vp]),
    "b200_spmv_sell_bcast_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _ll]),
    "b200_spmv_sell_halo_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _ll, _vp, _vp]),
    "b200_spmv_sell_ring_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _ll, _vp, _vp, _vp, _i, _u64]),
    "b200_minmax_i32": (_i, [_vp, _vp, _ll, C.POINTER(_i), C.POINTER(_i)]),
    "b200_used_column_blocks": (_i, [_vp, _vp, _ll, _i, _i, _vp]),
    "b200_ipc_get_handle": (_i, [_vp, _vp, _vp]),
    "b200_ipc_open_handle": (_i, [_vp, _vp, _vpp]),
    "b200_ipc_close_handle": (_i, [_vp, _vp]),
    "b200_comm_get_unique_id": (_i, [_vp]),
    "b200_comm_create": (_i, [_vp, _vp, _i, _i, _vpp]),
    "b200_comm_destroy": (_i, [_vp]),
    "b200_comm_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "b200_comm_check": (_i, [_vp]),
    "b200_comm_allreduce_sum_f64": (_i, [_vp, _vp, _ll]),
    "b200_comm_allgather_f64": (_i, [_vp, _vp, _ll]),
    "b200_comm_allgather_bytes": (_i, [_vp, _vp, _ll]),
    "b200_ctx_enable_peer_access": (_i, [_vp, _i]),
    "b200_halo_rows": (_i, [_vp, _vp, _i, _i, _ll, _ll, _vp, _vp]),
    "b200_mcast_supported": (_i, [_vp, C.POINTER(_i)]),
    "b200_mcast_create": (_i, [_vp, _i, _vpp, C.POINTER(_i)]),
    "b200_mcast_import_fd": (_i, [_vp, _i, _i, _vpp]),
    "b200_mcast_import_pid_fd": (_i, [_vp, _i, _i, _i, _vpp]),
    "b200_mcast_share": (_i, [_vp, _vp, _vpp]),
    "b200_mcast_add_device": (_i, [_vp]),
    "b200_mcast_bind": (_i, [_vp, _i]),
    "b200_mcast_pointers": (_i, [_vp, _vpp, _vpp, C.POINTER(_sz)]),
    "b200_mcast_step_buffers": (_i, [_vp, _vpp, _vpp]),
    "b200_mcast_allreduce_barrier": (_i, [_vp]),
    "b200_mcast_destroy": (_i, [_vp]),
    "b200_iterator_create": (_i, [_vp, _vp, C.POINTER(BlockF64), C.POINTER(IterDesc), _vpp]),
    "b200_iterator_run": (_i, [_vp, _i]),
    "b200_iterator_norm": (_i, [_vp, C.POINTER(C.c_double)]),
    "b200_iterator_state": (_i, [_vp, C.POINTER(C.c_ulonglong), _vpp, C.POINTER(C.c_ulonglong)]),
    "b200_iterator_destroy": (_i, [_vp]),
    "b200_sell_pack16_f64": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "b200_sell_pack16_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "b200_spmv_sell16_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    "b200_spmv_sell16_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    "b200_ell_to_csr_f64": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, C.POINTER(_ll)]),
    "b200_ell_to_csr_f32": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, C.POINTER(_ll)]),
    "b200_sell_to_csr_f64": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, C.POINTER(_ll)]),
    "b200_sell_to_csr_f32": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, C.POINTER(_ll)]),
    "b200_cmrs_to_csr_ptr": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "b200_format_advice": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(FormatAdvice)]),
    "b200_scale_f64": (_i, [_vp, _vp, _ll, _vp, _i]),
    "b200_sumsq_f64": (_i, [_vp, _vp, _ll, _vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libb200spmv.so (built in-tree by `make -C opencl-spmv-algorithms_b200 lib`)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with __graft_entry__.build(); "
                "there is no CPU fallback")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(status: int, where: str) -> None:
    if status != SUCCESS:
        raise B200Error(status, where, lib().b200_last_error().decode(errors="replace"))


def device_count() -> int:
    n = C.c_int(0)
    check(lib().b200_get_device_count(C.byref(n)), "b200_get_device_count")
    return n.value


_DT = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}


def suffix(dtype) -> str:
    return _DT[np.dtype(dtype)]


class DeviceArray:
    """A typed device allocation owned through the C ABI (b200_malloc / b200_free)."""

    def __init__(self, ctx: "Context", n: int, dtype):
        self.ctx, self.n, self.dtype = ctx, int(n), np.dtype(dtype)
        p = C.c_void_p()
        check(lib().b200_malloc(ctx.h, self.nbytes, C.byref(p)), "b200_malloc")
        self.ptr = p.value or 0
        self._p = p

    @classmethod
    def from_ptr(cls, ctx: "Context", ptr: int, n: int, dtype) -> "DeviceArray":
        """Non-owning view of device memory allocated elsewhere (e.g. a torch tensor's data_ptr())."""
        self = cls.__new__(cls)
        self.ctx, self.n, self.dtype, self.ptr, self._p = ctx, int(n), np.dtype(dtype), int(ptr), None
        self._borrowed = True
        return self

    @property
    def nbytes(self) -> int:
        return self.n * self.dtype.itemsize

    def upload(self, host: np.ndarray) -> "DeviceArray":
        host = np.ascontiguousarray(host, dtype=self.dtype)
        assert host.size == self.n, (host.size, self.n)
        self._keep = host  # the copy is asynchronous
        check(lib().b200_memcpy_h2d_async(self.ctx.h, self.ptr, host.ctypes.data, self.nbytes),
              "b200_memcpy_h2d_async")
        return self

    def download(self, count: int | None = None) -> np.ndarray:
        n = self.n if count is None else int(count)
        out = np.empty(n, self.dtype)
        check(lib().b200_memcpy_d2h(self.ctx.h, out.ctypes.data, self.ptr, n * self.dtype.itemsize),
              "b200_memcpy_d2h")
        return out

    def fill_bytes(self, value: int = 0) -> "DeviceArray":
        check(lib().b200_memset_async(self.ctx.h, self.ptr, value, self.nbytes), "b200_memset_async")
        return self

    def free(self) -> None:
        if self.ptr and not getattr(self, "_borrowed", False):
            check(lib().b200_free(self.ctx.h, self.ptr), "b200_free")
        self.ptr = 0

    def __del__(self):
        try:
            if self.ptr and self.ctx.h and not getattr(self, "_borrowed", False):
                lib().b200_free(self.ctx.h, self.ptr)
                self.ptr = 0
        except Exception:
            pass


class Context:
    """One device + one in-order stream (cl_context + cl_command_queue of the reference)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        h = C.c_void_p()
        if stream is None:
            check(lib().b200_ctx_create(device, C.byref(h)), "b200_ctx_create")
        else:
            check(lib().b200_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(h)),
                  "b200_ctx_create_on_stream")
        self.h = h
        self.device = device
        self._options = set()

    def set_option(self, name: str, value) -> None:
        """b200_ctx_set_option: override a tuning hook ("B200_CSR_LANES", ...) on this context;
        value None = back to automatic.  (The environment variables of the same names are read once,
        when the context is created.)"""
        check(lib().b200_ctx_set_option(self.h, name.encode(), None if value is None else str(value).encode()),
              "b200_ctx_set_option")
        (self._options.discard if value is None else self._options.add)(name)

    def clear_options(self) -> None:
        for name in list(self._options):
            self.set_option(name, None)

    def empty(self, n, dtype) -> DeviceArray:
        return DeviceArray(self, n, dtype)

    def zeros(self, n, dtype) -> DeviceArray:
        return DeviceArray(self, n, dtype).fill_bytes(0)

    def array(self, host: np.ndarray, dtype=None) -> DeviceArray:
        host = np.asarray(host)
        return DeviceArray(self, host.size, dtype or host.dtype).upload(host)

    def sync(self) -> None:
        check(lib().b200_sync(self.h), "b200_sync")

    def set_l2_persist(self, arr: DeviceArray | None) -> None:
        check(lib().b200_ctx_set_l2_persist(self.h, arr.ptr if arr else None,
                                            arr.nbytes if arr else 0), "b200_ctx_set_l2_persist")

    def set_launch_overlap(self, enable: bool) -> None:
        """b200_ctx_set_launch_overlap: SpMV launches become programmatic dependents (see the header
        for the contract: no un-synchronised writes to matrix arrays while enabled)."""
        check(lib().b200_ctx_set_launch_overlap(self.h, 1 if enable else 0), "b200_ctx_set_launch_overlap")

    def event(self) -> "Event":
        return Event(self)

    def record_graph(self) -> "GraphRecorder":
        """`with ctx.record_graph() as g: ...launches...` records them instead of running them;
        afterwards `g.launch()` replays the sequence with one call (b200_graph_*)."""
        return GraphRecorder(self)

    def close(self) -> None:
        if self.h:
            lib().b200_ctx_destroy(self.h)
            self.h = None


class GraphRecorder:
    """A recorded launch sequence (CUDA graph) on one context."""

    def __init__(self, ctx: Context):
        self.ctx, self.h = ctx, None

    def __enter__(self) -> "GraphRecorder":
        check(lib().b200_graph_begin(self.ctx.h), "b200_graph_begin")
        return self

    def __exit__(self, exc_type, exc, tb) -> None:
        h = C.c_void_p()
        status = lib().b200_graph_end(self.ctx.h, C.byref(h))  # always end the capture
        if exc_type is None:
            check(status, "b200_graph_end")
            self.h = h
        elif status == SUCCESS:
            lib().b200_graph_destroy(h)

    def launch(self) -> None:
        check(lib().b200_graph_launch(self.ctx.h, self.h), "b200_graph_launch")

    def __del__(self):
        try:
            if self.h:
                lib().b200_graph_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Event:
    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.h = C.c_void_p()
        check(lib().b200_event_create(ctx.h, C.byref(self.h)), "b200_event_create")

    def record(self) -> "Event":
        check(lib().b200_event_record(self.ctx.h, self.h), "b200_event_record")
        return self

    def elapsed_ms_until(self, stop: "Event") -> float:
        ms = C.c_float()
        check(lib().b200_event_elapsed_ms(self.h, stop.h, C.byref(ms)), "b200_event_elapsed_ms")
        return ms.value

    def __del__(self):
        try:
            if self.h:
                lib().b200_event_destroy(self.h)
                self.h = None
        except Exception:
            pass


from .formats import (CooMatrix, CsrMatrix, EllMatrix, EllCmMatrix, SellMatrix, Sell16Matrix, CmrsMatrix,  # noqa: E402,F401
                      CmrsPackedMatrix, algorithmic_bytes, build_all, partition_rows)
from .iterate import (Comm, Iterator, McastBlock, PeerBuffers, RowBlocks, equal_row_blocks, exchange_col_ranges,  # noqa: E402,F401
                      gpu_callables, halo_rows, power_iteration, power_iteration_ring, power_iteration_fused, x_upload_slices)
