#!/usr/bin/env python
"""Where does a multi-GPU power-iteration step spend its time?  Run under torchrun (one rank per GPU):
times, per rank and as the max over ranks, (a) the 256-byte all-reduce alone, replayed from a launch
graph and launch by launch, (b) the fused SELL kernel with its halo peer stores but no collective,
(c) kernel + all-reduce, on the bench's 7-point Laplacian block.  Prints one JSON object on rank 0."""
import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = load_package()
    L = pkg.lib()
    ctx = pkg.Context(local)
    nx = ny = 400
    nz = 50 * world
    n = nx * ny * nz
    blocks = pkg.equal_row_blocks(n, world)
    lo, hi = blocks.bounds(rank)
    nl = hi - lo
    nnz = L.b200_gen_laplace7_nnz(nx, ny, nz, lo, nl)
    rows, cols, vals = ctx.empty(nnz, np.int32), ctx.empty(nnz, np.int32), ctx.empty(nnz, np.float64)
    pkg.check(L.b200_gen_laplace7_coo(ctx.h, nx, ny, nz, lo, nl, rows.ptr, cols.ptr, vals.ptr), "gen")
    pkg.check(L.b200_offset_i32(ctx.h, rows.ptr, nnz, -lo), "rebase")
    coo = pkg.CooMatrix(ctx, nl, n, rows, cols, vals)
    sell = pkg.SellMatrix(pkg.CsrMatrix(coo), np.float64)
    comm = pkg.Comm(pkg, ctx, rank, world)
    bufs = pkg.PeerBuffers(pkg, ctx, blocks, rank, world)
    halo = pkg.halo_rows(pkg.exchange_col_ranges(pkg, ctx, coo.cols, lo, world), blocks, rank)
    pkg.check(L.b200_gen_uniform_f64(ctx.h, bufs.local[0].ptr, n, 11, 0.0, 1.0), "x0")
    acc = ctx.zeros(32, np.float64)
    lo_a, hi_a = (C.c_int * world)(*halo[0]), (C.c_int * world)(*halo[1])

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    acc2 = [ctx.zeros(32, np.float64), ctx.zeros(32, np.float64)]
    step_no = [0]

    def kernel(with_halo=True, cur=0, scale=None, sums=None):
        pkg.check(L.b200_spmv_sell_halo_f64(ctx.h, sell.data.ptr, sell.cols.ptr, bufs.local[cur].ptr, sell.row_indices.ptr, 32,
                                            sell.n_slices, nl, scale, (sums or acc).ptr,
                                            bufs.dst[1 - cur] if with_halo else (C.c_void_p * 1)(bufs.local[1 - cur].ptr),
                                            world if with_halo else 1, rank * blocks.count, lo_a if with_halo else None,
                                            hi_a if with_halo else None), "kernel")

    def with_memset():
        acc.fill_bytes(0)
        kernel()
        allreduce()

    def with_memset_scale():
        acc.fill_bytes(0)
        kernel(scale=acc2[0].ptr)
        allreduce()

    def true_step():
        k = step_no[0]
        step_no[0] += 1
        a = acc2[k % 2]
        a.fill_bytes(0)
        kernel(cur=k % 2, scale=acc2[(k - 1) % 2].ptr if k > 0 else None, sums=a)
        comm.allreduce_sum(a.ptr, 32)

    def true_step_fixed_buffers():
        k = step_no[0]
        step_no[0] += 1
        a = acc2[k % 2]
        a.fill_bytes(0)
        kernel(cur=0, scale=acc2[(k - 1) % 2].ptr if k > 0 else None, sums=a)
        comm.allreduce_sum(a.ptr, 32)

    def allreduce():
        comm.allreduce_sum(acc.ptr, 32)

    def timed(fn, reps, graph):
        for _ in range(3):
            fn()
        barrier()
        if graph:
            with ctx.record_graph() as g:
                for _ in range(reps):
                    fn()
            g.launch()
            barrier()
        a, b = ctx.event(), ctx.event()
        a.record()
        if graph:
            g.launch()
        else:
            for _ in range(reps):
                fn()
        b.record()
        barrier()
        return a.elapsed_ms_until(b) / reps * 1e3

    def both():
        kernel()
        allreduce()

    def sync_only():
        kernel(False)
        allreduce()

    out = {}
    for name, fn in (("allreduce_256B", allreduce), ("kernel_local_only", lambda: kernel(False)), ("kernel_with_halo_stores", kernel),
                     ("kernel_plus_allreduce", both), ("local_kernel_plus_allreduce", sync_only),
                     ("memset_kernel_allreduce", with_memset), ("memset_kernel_scaled_allreduce", with_memset_scale),
                     ("true_step_fixed_buffers", true_step_fixed_buffers), ("true_step", true_step)):
        for graph in (True, False):
            step_no[0] = 0
            pkg.check(L.b200_gen_uniform_f64(ctx.h, bufs.local[0].ptr, n, 11, 0.0, 1.0), "x0")
            pkg.check(L.b200_gen_uniform_f64(ctx.h, bufs.local[1].ptr, n, 11, 0.0, 1.0), "x0")
            acc2[0].upload(np.full(32, 1.0 / 32))
            acc2[1].upload(np.full(32, 1.0 / 32))
            us = timed(fn, 100, graph)
            t = torch.tensor([us], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[f"{name}{'_graph' if graph else '_direct'}_us"] = round(float(t[0]), 2)
    # the library's own iterator (b200_iterator_*), in this very process, under conditions bench.py adds one at a time
    def iterator_us(label, graph_steps=100, pre=None, post=None):
        pkg.check(L.b200_gen_uniform_f64(ctx.h, bufs.local[0].ptr, n, 11, 0.0, 1.0), "x0")
        bufs.local[1].fill_bytes(0)
        barrier()
        it = pkg.Iterator(pkg, ctx, comm, sell, blocks, rank, world, bufs.ptrs[:2], mode="fused", halo=halo,
                          graph_steps=graph_steps)
        it.run(1 + 2 * max(graph_steps, 2))
        barrier()
        if pre:
            pre()
        a, b = ctx.event(), ctx.event()
        a.record()
        it.run(100)
        b.record()
        barrier()
        if post:
            post()
        us = a.elapsed_ms_until(b) * 10.0
        it.close()
        t = torch.tensor([us], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[label] = round(float(t[0]), 2)

    iterator_us("iterator_graph100_us")
    iterator_us("iterator_graph10_us", graph_steps=10)
    iterator_us("iterator_direct_us", graph_steps=0)
    import bench
    holder = {}

    def sampler_on():
        holder["clk"] = bench.ClockSampler(local)
        holder["clk"].__enter__()

    def sampler_off():
        holder["clk"].__exit__()

    iterator_us("iterator_graph100_with_clock_sampler_us", pre=sampler_on, post=sampler_off)
    ctx2 = pkg.Context(local)
    big = ctx2.zeros(1 << 26, np.float64)
    iterator_us("iterator_graph100_second_context_us")
    ctx2.set_l2_persist(big)          # what the banded arm leaves behind: a device-wide persisting-L2 carve-out
    ctx2.sync()
    iterator_us("iterator_graph100_after_l2_persist_carveout_us")
    ctx2.set_l2_persist(None)
    ctx2.sync()
    iterator_us("iterator_graph100_after_clearing_it_us")
    bench.bind_to_gpu_numa(local)
    iterator_us("iterator_graph100_after_cpu_binding_us")
    # the NVSwitch-multicast all-reduce + barrier (b200_mcast_*): alone, and as the step's hand-over
    try:
        mc = pkg.McastBlock(pkg, ctx, rank, world)
    except pkg.B200Error as e:
        mc = None
        out["mcast"] = f"unavailable: {e}"
    if mc is not None:
        def mc_sync():
            pkg.check(L.b200_mcast_allreduce_barrier(mc.h), "mcast sync")
        for graph in (True, False):
            us = timed(mc_sync, 100, graph)
            t = torch.tensor([us], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[f"mcast_allreduce_barrier{'_graph' if graph else '_direct'}_us"] = round(float(t[0]), 2)
        pkg.check(L.b200_gen_uniform_f64(ctx.h, bufs.local[0].ptr, n, 11, 0.0, 1.0), "x0")
        bufs.local[1].fill_bytes(0)
        barrier()
        it = pkg.Iterator(pkg, ctx, None, sell, blocks, rank, world, bufs.ptrs[:2], mode="fused_mcast", halo=halo,
                          graph_steps=100, mcast=mc)
        it.run(201)
        barrier()
        a, b = ctx.event(), ctx.event()
        a.record()
        it.run(100)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_ms_until(b) * 10.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["iterator_mcast_graph100_us"] = round(float(t[0]), 2)
        it.close()
        barrier()
        mc.close()
    if rank == 0:
        print(json.dumps({"world": world, "nccl": comm.nccl_version(), **out}))
    barrier()
    bufs.close()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
