"""Iterated (power-iteration) SpMV over row-partitioned shards -- BASELINE.json configs[4].

The reference is single-device (it enumerates up to 8 GPUs but `break`s after the first,
csr.c:279); this mode is new.  Each rank owns a contiguous row block of A (global column
indices) and a full-length copy of x.  One step:

    y_r   = A_r . x                 local SpMV, written STRAIGHT INTO this rank's segment of the
                                    gather buffer (no pack step)
    s     = sum_r ||y_r||^2         1-element all-reduce
    y_r  *= 1/sqrt(s)               scale only the local segment (R/G elements, not R)
    x'    = all-gather(y_r)         in place: the segment is already where NCCL expects it

torch / torch.distributed are plumbing here (device buffers, the stream, NCCL over NVLink); the
arithmetic is three C-ABI calls (b200_spmv_*, b200_sumsq_f64, b200_scale_f64).  The step logic is
backend-agnostic -- it takes the three local operations as callables -- so the same code is
exercised on CPU with gloo in tests/test_distributed_cpu.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable


@dataclass
class RowBlocks:
    """Equal-count row blocks (what an all-gather needs): block r = [r*count, min((r+1)*count, n))."""
    n_rows: int
    world: int
    count: int  # rows per rank, a multiple of `align`

    def bounds(self, rank: int):
        lo = min(rank * self.count, self.n_rows)
        return lo, min(lo + self.count, self.n_rows)

    @property
    def padded(self) -> int:
        return self.count * self.world


def equal_row_blocks(n_rows: int, world: int, align: int = 32) -> RowBlocks:
    """Row blocks for matrices with (near-)uniform row lengths: ceil(n/world) rounded up to a
    multiple of `align` (= lcm of the SELL chunk 32 and the CMRS height 8), so every shard's
    chunks/strips coincide with those of the global build."""
    per = -(-n_rows // world)
    count = -(-per // align) * align
    return RowBlocks(n_rows, world, count)


@dataclass
class IterationResult:
    steps: int
    norm: float          # ||A x_{k-1}||_2 of the last step = dominant-eigenvalue estimate
    x: object            # the gathered, normalised vector (length blocks.padded)


def power_iteration(*, x_cur, x_next, rank: int, blocks: RowBlocks, steps: int,
                    spmv_local: Callable, sumsq: Callable, scale_inv_sqrt: Callable,
                    all_reduce_sum: Callable, all_gather_inplace: Callable,
                    on_step: Callable | None = None) -> IterationResult:
    """Run `steps` iterations.  x_cur / x_next are 1-D buffers of length blocks.padded (any type
    the callables understand; torch tensors in the product).

    spmv_local(x_full, y_segment)    y_segment[0:rows_of_rank] = A_r . x_full
    sumsq(y_segment, acc)            acc[0] += sum(y_segment**2)           (acc: 1-element buffer)
    scale_inv_sqrt(y_segment, acc)   y_segment *= 1/sqrt(acc[0])
    all_reduce_sum(acc)              in place over ranks
    all_gather_inplace(full, seg)    gather every rank's segment into `full` (seg is a view of it)
    """
    lo = rank * blocks.count
    norm2 = None
    for k in range(steps):
        seg = x_next[lo:lo + blocks.count]
        spmv_local(x_cur, seg)
        acc = sumsq(seg)
        all_reduce_sum(acc)
        scale_inv_sqrt(seg, acc)
        all_gather_inplace(x_next, seg)
        norm2 = acc
        x_cur, x_next = x_next, x_cur
        if on_step is not None:
            on_step(k)
    return IterationResult(steps, float(norm2[0]) ** 0.5 if norm2 is not None else float("nan"), x_cur)


def gpu_callables(pkg, ctx, matrix, n_rows_local: int):
    """The product wiring: three C-ABI kernels on `ctx`'s stream (create ctx on torch's current
    stream so NCCL ops and kernels are ordered), collectives via torch.distributed (NCCL)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    L = pkg.lib()
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")

    def view(t):
        return pkg.DeviceArray.from_ptr(ctx, t.data_ptr(), t.numel(), np.float64)

    def spmv_local(x_full, seg):
        matrix.spmv(view(x_full), view(seg[:n_rows_local]))

    def sumsq(seg):
        acc.zero_()
        pkg.check(L.b200_sumsq_f64(ctx.h, seg.data_ptr(), n_rows_local, acc.data_ptr()), "b200_sumsq_f64")
        return acc

    def scale_inv_sqrt(seg, a):
        pkg.check(L.b200_scale_f64(ctx.h, seg.data_ptr(), n_rows_local, a.data_ptr(), 1), "b200_scale_f64")

    def all_reduce_sum(a):
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(a, op=dist.ReduceOp.SUM)

    def all_gather_inplace(full, seg):
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_gather_into_tensor(full, seg)

    return dict(spmv_local=spmv_local, sumsq=sumsq, scale_inv_sqrt=scale_inv_sqrt,
                all_reduce_sum=all_reduce_sum, all_gather_inplace=all_gather_inplace)
