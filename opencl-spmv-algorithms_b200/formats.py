"""Host-side mirror of the reference drivers' flow, one class per format.

Each class owns the device arrays the corresponding reference driver uploads (same names, same
order: SURVEY.md section 8b) and exposes `spmv(x, y)`, which is one call into the C ABI with the
reference kernel's argument order.  Builds run on the GPU (csrc/build_formats.cu) from the device
COO triples, exactly as the C drivers in host/ do.  Nothing here touches the CPU oracle.

  reference driver      class         build                               launch
  coo.c                 CooMatrix     triples as parsed (coo.c:75-84)      b200_spmv_coo_*
  csr.c                 CsrMatrix     b200_build_csr_ptr (csr.c:72-91)     b200_spmv_csr_*
  ell.c                 EllMatrix     b200_build_ell_* (ell.c:68-164)      b200_spmv_ell_*
  (new layout)          EllCmMatrix   b200_build_ell_colmajor_*            b200_spmv_ellcm_*
  sigma_c.c             SellMatrix    b200_build_sell_* (sigma_c.c:71-202) b200_spmv_sell[64]_*
  cmrs.c                CmrsMatrix    b200_build_cmrs (cmrs.c:72-117)      b200_spmv_cmrs_*
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import (Context, CsrPlanInfo, DeviceArray, FormatAdvice, RowStats, check, lib, suffix)

I4 = 4


class CooMatrix:
    """Device COO triples in the order given (file order).  Values are kept in fp64 (what the
    drivers parse with %lg) and converted on the device for the fp32 path."""

    def __init__(self, ctx: Context, n_rows: int, n_cols: int, rows: DeviceArray, cols: DeviceArray,
                 vals: DeviceArray):
        self.ctx, self.n_rows, self.n_cols = ctx, int(n_rows), int(n_cols)
        self.rows, self.cols, self.vals64 = rows, cols, vals
        self.nnz = rows.n
        self._vals32 = None

    @classmethod
    def from_host(cls, ctx, n_rows, n_cols, rows, cols, vals):
        return cls(ctx, n_rows, n_cols, ctx.array(rows, np.int32), ctx.array(cols, np.int32),
                   ctx.array(vals, np.float64))

    def values(self, dtype) -> DeviceArray:
        if np.dtype(dtype) == np.float64:
            return self.vals64
        if self._vals32 is None:
            self._vals32 = self.ctx.empty(self.nnz, np.float32)
            check(lib().b200_convert_f64_to_f32(self.ctx.h, self.vals64.ptr, self._vals32.ptr,
                                                self.nnz), "b200_convert_f64_to_f32")
        return self._vals32

    def check_sorted(self) -> None:
        check(lib().b200_check_sorted_rows(self.ctx.h, self.rows.ptr, self.nnz, self.n_rows),
              "b200_check_sorted_rows")

    def spmv(self, x: DeviceArray, y: DeviceArray) -> None:
        v = self.values(x.dtype)
        fn = getattr(lib(), "b200_spmv_coo_" + suffix(x.dtype))
        check(fn(self.ctx.h, self.rows.ptr, self.cols.ptr, v.ptr, x.ptr, y.ptr, self.nnz,
                 self.n_rows), "b200_spmv_coo")

    def nbytes(self, dtype) -> int:
        V = np.dtype(dtype).itemsize
        return algorithmic_bytes("coo", V, n_rows=self.n_rows, n_cols=self.n_cols, nnz=self.nnz)


class CsrMatrix:
    def __init__(self, coo: CooMatrix, check_sorted: bool = True):
        self.ctx, self.coo = coo.ctx, coo
        self.n_rows, self.n_cols, self.nnz = coo.n_rows, coo.n_cols, coo.nnz
        if check_sorted:
            coo.check_sorted()
        self.ptr = self.ctx.empty(self.n_rows + 1, np.int32)
        check(lib().b200_build_csr_ptr(self.ctx.h, coo.rows.ptr, self.nnz, self.n_rows,
                                       self.ptr.ptr), "b200_build_csr_ptr")
        self.cols = coo.cols
        self._plan = None

    def plan(self):
        if self._plan is None:
            p = C.c_void_p()
            check(lib().b200_csr_plan_create(self.ctx.h, self.ptr.ptr, self.n_rows, C.byref(p)),
                  "b200_csr_plan_create")
            self._plan = p
        return self._plan

    def plan_info(self) -> CsrPlanInfo:
        info = CsrPlanInfo()
        check(lib().b200_csr_plan_get_info(self.plan(), C.byref(info)), "b200_csr_plan_get_info")
        return info

    def row_stats(self) -> RowStats:
        st = RowStats()
        check(lib().b200_row_length_stats(self.ctx.h, self.ptr.ptr, self.n_rows, C.byref(st)),
              "b200_row_length_stats")
        return st

    def spmv(self, x: DeviceArray, y: DeviceArray, use_plan: bool = True) -> None:
        v = self.coo.values(x.dtype)
        fn = getattr(lib(), "b200_spmv_csr_" + suffix(x.dtype))
        check(fn(self.ctx.h, self.ptr.ptr, self.cols.ptr, v.ptr, x.ptr, y.ptr, self.n_rows,
                 self.plan() if use_plan else None), "b200_spmv_csr")

    def nbytes(self, dtype) -> int:
        V = np.dtype(dtype).itemsize
        return algorithmic_bytes("csr", V, n_rows=self.n_rows, n_cols=self.n_cols, nnz=self.nnz)

    def advice(self, dtype) -> FormatAdvice:
        """b200_format_advice: per-format bytes for this matrix and the format to use."""
        a = FormatAdvice()
        check(lib().b200_format_advice(self.ctx.h, self.ptr.ptr, self.n_rows, self.n_cols, np.dtype(dtype).itemsize,
                                       C.byref(a)), "b200_format_advice")
        return a

    def __del__(self):
        try:
            if self._plan:
                lib().b200_csr_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass


class EllMatrix:
    """Row-major ELL: the reference's arrays (ell.c:118-164).  row_size defaults to the longest
    row (== the reference's longest_col whenever the last row is not the unique longest)."""

    def __init__(self, csr: CsrMatrix, dtype, row_size: int | None = None):
        self.ctx, self.csr, self.dtype = csr.ctx, csr, np.dtype(dtype)
        self.n_rows, self.n_cols, self.nnz = csr.n_rows, csr.n_cols, csr.nnz
        self.row_size = int(csr.row_stats().max_len if row_size is None else row_size)
        slots = self.n_rows * self.row_size
        self.cols = self.ctx.empty(slots, np.int32)
        self.data = self.ctx.empty(slots, self.dtype)
        fn = getattr(lib(), "b200_build_ell_" + suffix(self.dtype))
        check(fn(self.ctx.h, csr.ptr.ptr, csr.cols.ptr, csr.coo.vals64.ptr, self.n_rows,
                 self.row_size, self.cols.ptr, self.data.ptr), "b200_build_ell")

    def spmv(self, x: DeviceArray, y: DeviceArray) -> None:
        fn = getattr(lib(), "b200_spmv_ell_" + suffix(self.dtype))
        check(fn(self.ctx.h, self.data.ptr, self.cols.ptr, x.ptr, y.ptr, self.n_rows,
                 self.row_size), "b200_spmv_ell")

    def nbytes(self, dtype=None) -> int:
        return algorithmic_bytes("ell", self.dtype.itemsize, n_rows=self.n_rows,
                                 n_cols=self.n_cols, row_size=self.row_size)


class EllCmMatrix:
    """Column-major ELL device layout: transpose of the same K-padded matrix, pitch % 32 == 0."""

    def __init__(self, csr: CsrMatrix, dtype, row_size: int | None = None):
        self.ctx, self.csr, self.dtype = csr.ctx, csr, np.dtype(dtype)
        self.n_rows, self.n_cols, self.nnz = csr.n_rows, csr.n_cols, csr.nnz
        self.row_size = int(csr.row_stats().max_len if row_size is None else row_size)
        self.pitch = (self.n_rows + 31) // 32 * 32
        slots = self.pitch * self.row_size
        self.cols = self.ctx.empty(slots, np.int32)
        self.data = self.ctx.empty(slots, self.dtype)
        fn = getattr(lib(), "b200_build_ell_colmajor_" + suffix(self.dtype))
        check(fn(self.ctx.h, csr.ptr.ptr, csr.cols.ptr, csr.coo.vals64.ptr, self.n_rows,
                 self.row_size, self.pitch, self.cols.ptr, self.data.ptr),
              "b200_build_ell_colmajor")

    def spmv(self, x: DeviceArray, y: DeviceArray) -> None:
        fn = getattr(lib(), "b200_spmv_ellcm_" + suffix(self.dtype))
        check(fn(self.ctx.h, self.data.ptr, self.cols.ptr, x.ptr, y.ptr, self.n_rows,
                 self.row_size, self.pitch), "b200_spmv_ellcm")

    def nbytes(self, dtype=None) -> int:
        return algorithmic_bytes("ell", self.dtype.itemsize, n_rows=self.n_rows,
                                 n_cols=self.n_cols, row_size=self.row_size)


class SellMatrix:
    """SELL-32-sigma.  sigma <= 1 is the reference layout (sigma_c.c); `row_indices` (int32) is
    the reference's array, `slice_ptr` (int64) the overflow-safe one the kernels can also use."""

    def __init__(self, csr: CsrMatrix, dtype, sigma: int = 1, chunk: int = 32, wide: bool = False):
        self.ctx, self.csr, self.dtype = csr.ctx, csr, np.dtype(dtype)
        self.n_rows, self.n_cols, self.nnz = csr.n_rows, csr.n_cols, csr.nnz
        self.sigma, self.chunk = int(sigma), int(chunk)
        L = lib()
        self.n_slices = L.b200_sell_num_slices(self.n_rows, chunk)
        self.perm = self.ctx.empty(self.n_rows, np.int32) if sigma > 1 else None
        self.slice_ptr = self.ctx.empty(self.n_slices + 1, np.int64)
        total = C.c_longlong(0)
        check(L.b200_build_sell_ptr(self.ctx.h, csr.ptr.ptr, self.n_rows, chunk, self.sigma,
                                    self.perm.ptr if self.perm else None, self.slice_ptr.ptr,
                                    C.byref(total)), "b200_build_sell_ptr")
        self.total = total.value
        self.wide = bool(wide) or self.total > 0x7fffffff
        self.row_indices = None
        if not self.wide:
            self.row_indices = self.ctx.empty(self.n_slices + 1, np.int32)
            check(L.b200_sell_ptr_to_i32(self.ctx.h, self.slice_ptr.ptr, self.n_slices,
                                         self.row_indices.ptr), "b200_sell_ptr_to_i32")
        self.cols = self.ctx.empty(self.total, np.int32)
        self.data = self.ctx.empty(self.total, self.dtype)
        fn = getattr(L, "b200_build_sell_fill_" + suffix(self.dtype))
        check(fn(self.ctx.h, csr.ptr.ptr, csr.cols.ptr, csr.coo.vals64.ptr, self.n_rows, chunk,
                 self.perm.ptr if self.perm else None, self.slice_ptr.ptr, self.cols.ptr,
                 self.data.ptr), "b200_build_sell_fill")

    def plan(self):
        """Wide-chunk work list (only non-trivial on power-law inputs)."""
        if getattr(self, "_plan", None) is None:
            p = C.c_void_p()
            if self.wide:
                check(lib().b200_sell64_plan_create(self.ctx.h, self.slice_ptr.ptr, self.n_slices,
                                                    C.byref(p)), "b200_sell64_plan_create")
            else:
                check(lib().b200_sell_plan_create(self.ctx.h, self.row_indices.ptr, self.n_slices,
                                                  C.byref(p)), "b200_sell_plan_create")
            self._plan = p
        return self._plan

    def plan_extra_items(self) -> int:
        n = C.c_int(0)
        check(lib().b200_sell_plan_extra_items(self.plan(), C.byref(n)), "b200_sell_plan_extra_items")
        return n.value

    def spmv(self, x: DeviceArray, y: DeviceArray, n_out: int | None = None,
             use_plan: bool = True) -> None:
        n_out = self.n_rows if n_out is None else n_out
        perm = self.perm.ptr if self.perm else None
        if self.wide:
            fn = getattr(lib(), "b200_spmv_sell64_" + suffix(self.dtype))
            p = self.slice_ptr.ptr
        else:
            fn = getattr(lib(), "b200_spmv_sell_" + suffix(self.dtype))
            p = self.row_indices.ptr
        check(fn(self.ctx.h, self.data.ptr, self.cols.ptr, x.ptr, y.ptr, p, self.chunk,
                 self.n_slices, n_out, perm, self.plan() if use_plan else None), "b200_spmv_sell")

    def __del__(self):
        try:
            if getattr(self, "_plan", None):
                lib().b200_sell_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    def nbytes(self, dtype=None) -> int:
        return algorithmic_bytes("sell", self.dtype.itemsize, n_rows=self.n_rows,
                                 n_cols=self.n_cols, padded=self.total, n_slices=self.n_slices,
                                 ptr_bytes=8 if self.wide else 4, perm=self.perm is not None)


class Sell16Matrix:
    """SELL-32 with 16-bit column deltas (b200_sell_pack16_*): a derived layout of a sigma = 1
    SellMatrix, 2 + V bytes per entry.  Raises B200Error(UNSUPPORTED) when a chunk spans > 65536 columns."""

    def __init__(self, sell: "SellMatrix"):
        assert sell.perm is None and sell.row_indices is not None, "sell16: sigma = 1, int32 chunk pointers"
        self.ctx, self.sell, self.dtype = sell.ctx, sell, sell.dtype
        self.n_rows, self.n_cols, self.nnz = sell.n_rows, sell.n_cols, sell.nnz
        self.chunk_base = self.ctx.empty(sell.n_slices, np.int32)
        self.delta16 = self.ctx.empty(sell.total, np.uint16)
        fn = getattr(lib(), "b200_sell_pack16_" + suffix(self.dtype))
        check(fn(self.ctx.h, sell.data.ptr, sell.cols.ptr, sell.row_indices.ptr, sell.n_slices, self.n_cols,
                 self.chunk_base.ptr, self.delta16.ptr), "b200_sell_pack16")

    def spmv(self, x: DeviceArray, y: DeviceArray, n_out: int | None = None) -> None:
        s = self.sell
        fn = getattr(lib(), "b200_spmv_sell16_" + suffix(self.dtype))
        check(fn(self.ctx.h, s.data.ptr, self.delta16.ptr, self.chunk_base.ptr, x.ptr, y.ptr, s.row_indices.ptr, 32,
                 s.n_slices, self.n_rows if n_out is None else n_out, self.n_cols), "b200_spmv_sell16")

    def nbytes(self, dtype=None) -> int:
        V, s = self.dtype.itemsize, self.sell
        return s.total * (2 + V) + (s.n_slices + 1) * I4 + s.n_slices * I4 + (self.n_cols + self.n_rows) * V


class CmrsMatrix:
    def __init__(self, csr: CsrMatrix, height: int = 8):
        self.ctx, self.csr, self.height = csr.ctx, csr, int(height)
        self.n_rows, self.n_cols, self.nnz = csr.n_rows, csr.n_cols, csr.nnz
        L = lib()
        self.n_strips = L.b200_cmrs_num_strips(self.n_rows, height)
        self.strip_ptr = self.ctx.empty(self.n_strips + 1, np.int32)
        self.row_in_strip = self.ctx.empty(self.nnz, np.int32)
        check(L.b200_build_cmrs(self.ctx.h, csr.coo.rows.ptr, csr.ptr.ptr, self.nnz, self.n_rows,
                                height, self.strip_ptr.ptr, self.row_in_strip.ptr),
              "b200_build_cmrs")
        self.cols = csr.cols

    def plan(self):
        """Long-strip work list (only non-trivial on power-law inputs)."""
        if getattr(self, "_plan", None) is None:
            p = C.c_void_p()
            check(lib().b200_cmrs_plan_create(self.ctx.h, self.strip_ptr.ptr, self.n_strips, C.byref(p)),
                  "b200_cmrs_plan_create")
            self._plan = p
        return self._plan

    def plan_extra_items(self) -> int:
        n = C.c_int(0)
        check(lib().b200_cmrs_plan_extra_items(self.plan(), C.byref(n)), "b200_cmrs_plan_extra_items")
        return n.value

    def plan_stream_tiles(self) -> int:
        """> 0: the plan found skewed strip lengths and selected the nnz-split kernel."""
        n = C.c_int(0)
        check(lib().b200_cmrs_plan_stream_tiles(self.plan(), C.byref(n)), "b200_cmrs_plan_stream_tiles")
        return n.value

    def spmv(self, x: DeviceArray, y: DeviceArray, use_plan: bool = True) -> None:
        v = self.csr.coo.values(x.dtype)
        fn = getattr(lib(), "b200_spmv_cmrs_" + suffix(x.dtype))
        check(fn(self.ctx.h, v.ptr, self.cols.ptr, self.strip_ptr.ptr, self.row_in_strip.ptr,
                 x.ptr, y.ptr, self.n_strips, self.height, self.n_rows,
                 self.plan() if use_plan else None), "b200_spmv_cmrs")

    def packed(self) -> "CmrsPackedMatrix":
        return CmrsPackedMatrix(self)

    def __del__(self):
        try:
            if getattr(self, "_plan", None):
                lib().b200_cmrs_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    def nbytes(self, dtype) -> int:
        V = np.dtype(dtype).itemsize
        return algorithmic_bytes("cmrs", V, n_rows=self.n_rows, n_cols=self.n_cols, nnz=self.nnz,
                                 n_strips=self.n_strips)


class CmrsPackedMatrix:
    """CMRS with row_in_strip folded into the top 5 bits of the column word (b200_cmrs_pack): a
    derived device layout, 4 + V bytes per entry; the reference arrays stay in `cmrs`."""

    def __init__(self, cmrs: CmrsMatrix):
        self.ctx, self.cmrs = cmrs.ctx, cmrs
        self.n_rows, self.n_cols, self.nnz = cmrs.n_rows, cmrs.n_cols, cmrs.nnz
        self.packed = self.ctx.empty(self.nnz, np.int32)
        check(lib().b200_cmrs_pack(self.ctx.h, cmrs.cols.ptr, cmrs.row_in_strip.ptr, self.nnz,
                                   self.n_cols, cmrs.height, self.packed.ptr), "b200_cmrs_pack")

    def spmv(self, x: DeviceArray, y: DeviceArray, use_plan: bool = True) -> None:
        c = self.cmrs
        v = c.csr.coo.values(x.dtype)
        fn = getattr(lib(), "b200_spmv_cmrs_packed_" + suffix(x.dtype))
        check(fn(self.ctx.h, v.ptr, self.packed.ptr, c.strip_ptr.ptr, x.ptr, y.ptr, c.n_strips,
                 c.height, c.n_rows, c.plan() if use_plan else None), "b200_spmv_cmrs_packed")

    def nbytes(self, dtype) -> int:
        V = np.dtype(dtype).itemsize
        return self.nnz * (I4 + V) + (self.cmrs.n_strips + 1) * I4 + (self.n_cols + self.n_rows) * V


def algorithmic_bytes(fmt: str, V: int, *, n_rows: int, n_cols: int, nnz: int = 0,
                      row_size: int = 0, padded: int = 0, n_slices: int = 0, ptr_bytes: int = 4,
                      perm: bool = False, n_strips: int = 0) -> int:
    """True bytes one SpMV must move: values + indices + pointers, x once,
    y once.  I = 4.  COO's zero-fill of y is not counted."""
    xy = n_cols * V + n_rows * V
    if fmt == "coo":
        return nnz * (2 * I4 + V) + xy
    if fmt == "csr":
        return nnz * (I4 + V) + (n_rows + 1) * I4 + xy
    if fmt in ("ell", "ellcm"):
        return n_rows * row_size * (I4 + V) + xy
    if fmt == "sell":
        return padded * (I4 + V) + (n_slices + 1) * ptr_bytes + (n_rows * I4 if perm else 0) + xy
    if fmt == "cmrs":
        return nnz * (2 * I4 + V) + (n_strips + 1) * I4 + xy
    raise ValueError(fmt)


def build_all(coo_sorted: CooMatrix, dtype, coo_any: CooMatrix | None = None, sigma: int = 1,
              ell: bool = True):
    """The five formats of the reference from one row-sorted COO (COO itself may be given in a
    different order, as coo.c reads the column-major cant.mtx)."""
    csr = CsrMatrix(coo_sorted)
    out = {"coo": coo_any or coo_sorted, "csr": csr}
    if ell:
        out["ell"] = EllMatrix(csr, dtype)
        out["ellcm"] = EllCmMatrix(csr, dtype)
    out["sell"] = SellMatrix(csr, dtype, sigma=sigma)
    out["cmrs"] = CmrsMatrix(csr)
    return out


def partition_rows(ptr_host: np.ndarray, n_parts: int, align: int = 32) -> np.ndarray:
    """nnz-balanced contiguous row blocks, cut points multiples of `align` (host function)."""
    ptr_host = np.ascontiguousarray(ptr_host, np.int32)
    cuts = np.empty(n_parts + 1, np.int32)
    check(lib().b200_partition_rows(ptr_host.ctypes.data, len(ptr_host) - 1, n_parts, align,
                                    cuts.ctypes.data), "b200_partition_rows")
    return cuts
