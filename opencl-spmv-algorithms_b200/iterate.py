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


def x_upload_slices(n_cols: int, world: int, itemsize: int) -> tuple[int, int]:
    """(entries per rank, padded length) for the sharded upload of a replicated x: rank r uploads entries
    [r*per, (r+1)*per) from the host and the ranks all-gather the slices over NVLink
    (b200_comm_allgather_bytes).  `per` = ceil(n_cols / world) rounded up so that a slice is a whole number
    of 16-byte units (every rank's segment of the gathered buffer stays 16-byte aligned); entries of the
    last slices beyond n_cols are padding."""
    per = -(-n_cols // world)
    per = -(-per * itemsize // 16) * 16 // itemsize
    return per, per * world


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


# ---------------------------------------------------------------------------------------------
# The product path: communicator + iterator live in the library (csrc/iterate.cu); Python only
# passes pointers.  The step loops further down are kept as the readable mirror the tests compare
# the library against.
# ---------------------------------------------------------------------------------------------
class Comm:
    """b200_comm: an NCCL communicator created by the library on a context's stream.  `exchange`
    carries rank 0's 128-byte id to the other ranks: a callable bytes -> bytes (identity on one
    rank); with torch.distributed initialised the default broadcasts it."""

    def __init__(self, pkg, ctx, rank: int, world: int, exchange: Callable | None = None):
        import ctypes as C
        self.pkg, self.ctx, self.rank, self.world = pkg, ctx, rank, world
        L = pkg.lib()
        ident = (C.c_ubyte * pkg.COMM_ID_BYTES)()
        if rank == 0:
            pkg.check(L.b200_comm_get_unique_id(ident), "b200_comm_get_unique_id")
        if exchange is None:
            def exchange(b):
                import torch.distributed as dist
                box = [b]
                dist.broadcast_object_list(box, src=0)
                return box[0]
        got = exchange(bytes(ident))
        ident = (C.c_ubyte * pkg.COMM_ID_BYTES).from_buffer_copy(got)
        self.h = C.c_void_p()
        pkg.check(L.b200_comm_create(ctx.h, ident, rank, world, C.byref(self.h)), "b200_comm_create")

    def nccl_version(self) -> int:
        import ctypes as C
        v = C.c_int(0)
        self.pkg.check(self.pkg.lib().b200_comm_info(self.h, None, None, C.byref(v)), "b200_comm_info")
        return v.value

    def check(self) -> None:
        """Poll for an asynchronous NCCL error (a dead peer, a failed link): raises B200Error."""
        self.pkg.check(self.pkg.lib().b200_comm_check(self.h), "b200_comm_check")

    def allreduce_sum(self, buf_ptr: int, count: int) -> None:
        self.pkg.check(self.pkg.lib().b200_comm_allreduce_sum_f64(self.h, buf_ptr, count), "b200_comm_allreduce_sum_f64")

    def allgather(self, full_ptr: int, count_per_rank: int) -> None:
        self.pkg.check(self.pkg.lib().b200_comm_allgather_f64(self.h, full_ptr, count_per_rank), "b200_comm_allgather_f64")

    def allgather_bytes(self, full_ptr: int, bytes_per_rank: int) -> None:
        self.pkg.check(self.pkg.lib().b200_comm_allgather_bytes(self.h, full_ptr, bytes_per_rank), "b200_comm_allgather_bytes")

    def close(self) -> None:
        if self.h:
            self.pkg.lib().b200_comm_destroy(self.h)
            self.h = None


class McastBlock:
    """b200_mcast: this rank's share of an NVSwitch multicast block (all-reduce + barrier of the iterated
    mode through multimem.red).  Collective constructor over `world` processes: rank 0 creates the
    object and exports a file descriptor, the others fetch it with pidfd_getfd; `all_gather_object` and
    `barrier` carry the set-up (torch.distributed by default).  Raises B200Error(UNSUPPORTED) on every
    rank alike when the box cannot do it."""

    def __init__(self, pkg, ctx, rank: int, world: int, all_gather_object=None, barrier=None):
        import ctypes as C
        import os
        self.pkg, self.ctx, self.h = pkg, ctx, None
        L = pkg.lib()
        if all_gather_object is None or barrier is None:
            import torch.distributed as dist
            all_gather_object = all_gather_object or dist.all_gather_object
            barrier = barrier or dist.barrier

        def agree(status, where):
            """every rank learns whether all ranks succeeded, so that nobody waits for a rank that gave up"""
            box = [None] * world
            all_gather_object(box, (status, pkg.lib().b200_last_error().decode(errors="replace") if status else ""))
            bad = [(r, st, msg) for r, (st, msg) in enumerate(box) if st]
            if bad:
                self.close()
                raise pkg.B200Error(bad[0][1], where, f"rank {bad[0][0]}: {bad[0][2]}")

        h, fd = C.c_void_p(), C.c_int(-1)
        status = L.b200_mcast_create(ctx.h, world, C.byref(h), C.byref(fd)) if rank == 0 else 0
        if rank == 0 and status == 0:
            self.h = h
        box = [None] * world
        all_gather_object(box, (os.getpid(), fd.value, status))
        owner_pid, owner_fd, st0 = box[0]
        if st0 == 0 and rank != 0:
            status = L.b200_mcast_import_pid_fd(ctx.h, world, owner_pid, owner_fd, C.byref(h))
            if status == 0:
                self.h = h
        agree(status or st0, "b200_mcast_create / import")
        agree(L.b200_mcast_add_device(self.h), "b200_mcast_add_device")
        barrier()
        agree(L.b200_mcast_bind(self.h, rank), "b200_mcast_bind")
        barrier()

    def close(self) -> None:
        if self.h:
            self.pkg.lib().b200_mcast_destroy(self.h)
            self.h = None


class Iterator:
    """b200_iterator: the power iteration run by the library (launch-graph replay, NCCL from C).

    matrix: SellMatrix (sigma = 1, int32 pointers) or CsrMatrix of this rank's row block.
    x_ptrs: x_ptrs[b][r] = device pointer of x buffer b of rank r as this process sees it
            (PeerBuffers.ptrs[:2] for mode 'fused'; for 'allgather' only [b][rank] is read)."""

    def __init__(self, pkg, ctx, comm: Comm | None, matrix, blocks: RowBlocks, rank: int, world: int,
                 x_ptrs, mode: str = "fused", halo=None, graph_steps: int = 10, mcast: "McastBlock | None" = None):
        import ctypes as C
        self.pkg, self.ctx, self.matrix, self.comm = pkg, ctx, matrix, comm
        L = pkg.lib()
        n_local = blocks.bounds(rank)[1] - blocks.bounds(rank)[0]
        blk = pkg.BlockF64()
        if hasattr(matrix, "row_indices"):
            assert matrix.perm is None and matrix.row_indices is not None, "SELL-32, sigma = 1, int32 pointers"
            blk.format, blk.n_slices = pkg.FORMAT_SELL, matrix.n_slices
            blk.ptr, blk.indices, blk.data = matrix.row_indices.ptr, matrix.cols.ptr, matrix.data.ptr
        else:
            blk.format, blk.n_slices = pkg.FORMAT_CSR, 0
            blk.ptr, blk.indices, blk.data = matrix.ptr.ptr, matrix.cols.ptr, matrix.coo.values("f8").ptr
            blk.csr_plan = matrix.plan()
        blk.n_rows = n_local
        d = pkg.IterDesc()
        d.mode = {"fused": pkg.ITER_FUSED, "allgather": pkg.ITER_ALLGATHER, "fused_mcast": pkg.ITER_FUSED_MCAST}[mode]
        if mode == "fused_mcast":
            d.mcast, comm = mcast.h, None
        d.world, d.rank, d.rows_per_rank, d.graph_steps = world, rank, blocks.count, graph_steps
        self._tables = [(C.c_void_p * world)(*[int(p) if p else None for p in x_ptrs[b]]) for b in range(2)]
        d.x[0] = C.cast(self._tables[0], C.POINTER(C.c_void_p))
        d.x[1] = C.cast(self._tables[1], C.POINTER(C.c_void_p))
        if halo is not None:
            self._lo, self._hi = (C.c_int * world)(*halo[0]), (C.c_int * world)(*halo[1])
            d.halo_lo, d.halo_hi = C.cast(self._lo, C.POINTER(C.c_int)), C.cast(self._hi, C.POINTER(C.c_int))
        self.h = C.c_void_p()
        pkg.check(L.b200_iterator_create(ctx.h, comm.h if comm else None, C.byref(blk), C.byref(d), C.byref(self.h)),
                  "b200_iterator_create")

    def run(self, steps: int) -> None:
        self.pkg.check(self.pkg.lib().b200_iterator_run(self.h, steps), "b200_iterator_run")

    def norm(self) -> float:
        import ctypes as C
        v = C.c_double(0.0)
        self.pkg.check(self.pkg.lib().b200_iterator_norm(self.h, C.byref(v)), "b200_iterator_norm")
        return v.value

    def state(self):
        """(steps issued, device pointer of the current x buffer, kernels + collectives issued)."""
        import ctypes as C
        k, x, n = C.c_ulonglong(0), C.c_void_p(), C.c_ulonglong(0)
        self.pkg.check(self.pkg.lib().b200_iterator_state(self.h, C.byref(k), C.byref(x), C.byref(n)), "b200_iterator_state")
        return k.value, x.value, n.value

    def close(self) -> None:
        if self.h:
            self.pkg.lib().b200_iterator_destroy(self.h)
            self.h = None


# ---------------------------------------------------------------------------------------------
# Fused variant: the SpMV kernel itself stores its y-block into every rank's next-x buffer over
# NVLink (peer memory mapped with CUDA IPC) -- no all-gather, no pack, no separate scale pass.
# ---------------------------------------------------------------------------------------------
class PeerBuffers:
    """Two full-length x buffers (and one small sync block) per rank, allocated through the C ABI and
    mapped on every rank (b200_ipc_get_handle / b200_ipc_open_handle; handles travel with
    all_gather_object)."""
    SYNC_BLOCK_BYTES = 16384  # B200_SYNC_BLOCK_BYTES

    def __init__(self, pkg, ctx, blocks: RowBlocks, rank: int, world: int, all_gather_object=None):
        import ctypes as C
        import numpy as np
        self.pkg, self.ctx, self.rank, self.world = pkg, ctx, rank, world
        L = pkg.lib()
        self.local = [ctx.zeros(blocks.padded, np.float64) for _ in range(2)]
        # sync block of the ring exchange (b200_spmv_sell_ring_f64): flags + partial sums, zero-filled
        self.sync_block = ctx.zeros(self.SYNC_BLOCK_BYTES // 8, np.float64)
        self.ring_step = 0  # steps issued so far with the ring kernel (flags only ever grow)
        ctx.sync()
        handles = []
        for buf in self.local + [self.sync_block]:
            h = (C.c_ubyte * 64)()
            pkg.check(L.b200_ipc_get_handle(ctx.h, buf.ptr, h), "b200_ipc_get_handle")
            handles.append(bytes(h))
        gathered = [handles]
        if world > 1:
            if all_gather_object is None:
                import torch.distributed as dist
                all_gather_object = dist.all_gather_object
            gathered = [None] * world
            all_gather_object(gathered, handles)
        self._opened = []
        self.ptrs = []  # ptrs[b][r] = device pointer of buffer b of rank r, valid in THIS process
        for b in range(3):
            row = []
            for r in range(world):
                if r == rank:
                    row.append((self.local + [self.sync_block])[b].ptr)
                else:
                    p = C.c_void_p()
                    hb = (C.c_ubyte * 64).from_buffer_copy(gathered[r][b])
                    pkg.check(L.b200_ipc_open_handle(ctx.h, hb, C.byref(p)), "b200_ipc_open_handle")
                    self._opened.append(p.value)
                    row.append(p.value)
            self.ptrs.append(row)
        # (rotating the destination order by rank was tried at 8 GPUs: no gain, 1.85 vs 1.55 ms/step
        # on another box -- within box-to-box noise -- so the plain order is kept)
        self.dst = [(C.c_void_p * world)(*row) for row in self.ptrs[:2]]
        self.sync = (C.c_void_p * world)(*self.ptrs[2])  # sync block of every rank, by rank

    def close(self):
        L = self.pkg.lib()
        self.ctx.sync()
        for p in self._opened:
            L.b200_ipc_close_handle(self.ctx.h, p)
        self._opened = []


def halo_rows(col_ranges, blocks: RowBlocks, rank: int):
    """Which rows of rank `rank`'s block each rank reads as columns.

    col_ranges[d] = (cmin_d, cmax_d): smallest / largest column index in rank d's matrix block
    (b200_minmax_i32 over its column array, exchanged once with all_gather).  Returns (lo, hi): two
    lists of LOCAL row numbers, destination d needs rows [lo[d], hi[d]) of this rank's y-block; the
    rank itself always gets its whole block (it is the next x of its own rows AND the result).
    Pure host function (also exercised on CPU in tests/test_distributed_cpu.py)."""
    b0, b1 = blocks.bounds(rank)
    lo, hi = [], []
    for d, (cmin, cmax) in enumerate(col_ranges):
        if d == rank:
            lo.append(0)
            hi.append(b1 - b0)
            continue
        first, last = max(int(cmin), b0), min(int(cmax) + 1, b1)
        if last <= first:
            lo.append(0)
            hi.append(0)
        else:
            lo.append(first - b0)
            hi.append(last - b0)
    return lo, hi


def exchange_col_ranges(pkg, ctx, cols, row_begin: int, world: int):
    """(cmin, cmax) of every rank's column array `cols` (device, GLOBAL column indices): one tiny
    reduction on the device + one all_gather of two ints."""
    import ctypes as C
    lo, hi = C.c_int(0), C.c_int(0)
    pkg.check(pkg.lib().b200_minmax_i32(ctx.h, cols.ptr, cols.n, C.byref(lo), C.byref(hi)), "b200_minmax_i32")
    mine = (lo.value, hi.value)
    if world == 1:
        return [mine]
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, mine)
    return out


def power_iteration_ring(pkg, ctx, sell, bufs: PeerBuffers, rank: int, blocks: RowBlocks, steps: int,
                         halo=None) -> IterationResult:
    """`steps` power-iteration steps with NO collective call inside the loop: one kernel launch per
    step (b200_spmv_sell_ring_f64).  The kernel waits on the peers' "previous step done" flags, takes
    1/||x|| from the partial sums they left in this rank's sync block, runs the SpMV + halo stores and
    releases its own flag.  bufs.local[bufs.ring_step % 2] holds the current vector; consecutive calls
    continue the run (the step counter lives in `bufs`).  The result's norm is read after a barrier
    over all ranks (the peers' last partial sums must have landed)."""
    import numpy as np
    L = pkg.lib()
    n_local = blocks.bounds(rank)[1] - blocks.bounds(rank)[0]
    assert sell.perm is None and sell.row_indices is not None, "ring path: SELL-32, sigma = 1, int32 pointers"
    row_lo = row_hi = None
    if halo is not None:
        import ctypes as C
        row_lo = (C.c_int * bufs.world)(*halo[0])
        row_hi = (C.c_int * bufs.world)(*halo[1])
    for _ in range(steps):
        k = bufs.ring_step
        pkg.check(L.b200_spmv_sell_ring_f64(ctx.h, sell.data.ptr, sell.cols.ptr, bufs.local[k % 2].ptr,
                                            sell.row_indices.ptr, 32, sell.n_slices, n_local, bufs.dst[(k + 1) % 2],
                                            bufs.world, rank * blocks.count, row_lo, row_hi, bufs.sync, rank, k),
                  "b200_spmv_sell_ring_f64")
        bufs.ring_step = k + 1
    ctx.sync()
    if bufs.world > 1:
        import torch.distributed as dist
        dist.barrier()
    last = bufs.ring_step - 1
    words = bufs.sync_block.download()
    at = 16 + (last & 1) * 16 * 32
    norm2 = float(np.sum(words[at:at + bufs.world * 32])) if last >= 0 else float("nan")
    return IterationResult(steps, norm2 ** 0.5, bufs.local[bufs.ring_step % 2])


def power_iteration_fused(pkg, ctx, sell, bufs: PeerBuffers, rank: int, blocks: RowBlocks, steps: int,
                          first_step: int = 0, acc=None, halo=None) -> IterationResult:
    """`steps` power-iteration steps with the fused SpMV + exchange kernel.  bufs.local[first_step % 2]
    holds the current (unnormalised) vector on every rank; `acc` carries the two ||y||^2 scalars
    between calls (pass the returned result's .acc back in to continue a run).

    halo = (lo, hi) from halo_rows(): every destination receives only the rows it reads (its own block
    in full); the buffers then hold a rank's own block plus its halo, not the whole vector.
    halo = None: every row goes to every rank (the buffers hold the whole vector, as an all-gather)."""
    import torch
    import torch.distributed as dist
    L = pkg.lib()
    n_local = blocks.bounds(rank)[1] - blocks.bounds(rank)[0]
    slots = 32  # B200_SUMSQ_SLOTS partial sums of ||y||^2, added up by the consuming kernel
    if acc is None:
        acc = [torch.zeros(slots, dtype=torch.float64, device="cuda") for _ in range(2)]
    assert sell.perm is None and sell.row_indices is not None, "fused path: SELL-32, sigma = 1, int32 pointers"
    row_lo = row_hi = None
    if halo is not None:
        import ctypes as C
        assert len(halo[0]) == bufs.world and len(halo[1]) == bufs.world
        row_lo = (C.c_int * bufs.world)(*halo[0])
        row_hi = (C.c_int * bufs.world)(*halo[1])

    def step(k):
        """One launch + one tiny all-reduce: memset of the 32 slots, the fused kernel, NCCL."""
        cur, nxt = k % 2, (k + 1) % 2
        scale = acc[(k - 1) % 2].data_ptr() if k > 0 else None
        a = acc[k % 2]
        pkg.check(L.b200_memset_async(ctx.h, a.data_ptr(), 0, 8 * slots), "b200_memset_async")
        pkg.check(L.b200_spmv_sell_halo_f64(ctx.h, sell.data.ptr, sell.cols.ptr, bufs.local[cur].ptr,
                                            sell.row_indices.ptr, 32, sell.n_slices, n_local, scale,
                                            a.data_ptr(), bufs.dst[nxt], bufs.world, rank * blocks.count,
                                            row_lo, row_hi), "b200_spmv_sell_halo_f64")
        if bufs.world > 1:
            dist.all_reduce(a, op=dist.ReduceOp.SUM)  # the norm AND the barrier that orders peer writes

    k, end = first_step, first_step + steps
    while k < end:
        step(k)
        k += 1
    last = end - 1
    res = IterationResult(steps, float(acc[last % 2].sum()) ** 0.5, bufs.local[(last + 1) % 2])
    res.acc = acc
    res.next_step = end
    return res
