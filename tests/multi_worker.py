"""One rank of the real multi-GPU parity test (tests/test_gpu_multi.py starts `world` of these, one
per GPU).  Every formulation of the iterated mode runs across REAL devices -- CUDA IPC peer mappings,
the library's own NCCL communicator -- on a small 7-point Laplacian and is compared with the CPU
oracle's power iteration; the row-sharded single-shot SpMV of every format is compared with the
oracle as well.  gloo (CPU) carries the set-up objects only: the unique id, the IPC handles, the
column ranges.

    python tests/multi_worker.py RANK WORLD PORT OUT_JSON
"""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from __graft_entry__ import load_package
    from oracle import binding as O
    from test_distributed_cpu import laplace7
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = load_package()
    L = pkg.lib()
    ctx = pkg.Context(rank)
    res = {"rank": rank, "checks": {}}

    def check(name, ok, detail=""):
        res["checks"][name] = {"ok": bool(ok), "detail": str(detail)}

    nx, ny, nz, steps = 20, 16, 8 * world + 3, 31
    n, rows, cols, vals = laplace7(nx, ny, nz)
    blocks = pkg.equal_row_blocks(n, world)
    x0 = np.zeros(blocks.padded)
    x0[:n] = np.random.default_rng(3).uniform(0, 1, n)
    ptr, _ = O.build_csr(n, rows)
    x = x0[:n].copy()
    for _ in range(steps):                      # the oracle: CPU power iteration
        y = O.spmv_csr(n, ptr, cols, vals, x)
        nrm = float(np.linalg.norm(y))
        x = y / nrm
    b0, b1 = blocks.bounds(rank)
    sel = slice(ptr[b0], ptr[b1])
    coo = pkg.CooMatrix.from_host(ctx, b1 - b0, n, rows[sel] - b0, cols[sel], vals[sel])
    csr = pkg.CsrMatrix(coo)
    csr.plan()
    sell = pkg.SellMatrix(csr, np.float64)

    comm = pkg.Comm(pkg, ctx, rank, world)      # NCCL communicator made by the library; id over gloo
    res["nccl_version"] = comm.nccl_version()
    bufs = pkg.PeerBuffers(pkg, ctx, blocks, rank, world, all_gather_object=dist.all_gather_object)
    ranges = pkg.exchange_col_ranges(pkg, ctx, coo.cols, b0, world)
    halo = pkg.halo_rows(ranges, blocks, rank)
    # the C-ABI twin of halo_rows
    import ctypes as C
    cmin = (C.c_int * world)(*[r[0] for r in ranges])
    cmax = (C.c_int * world)(*[r[1] for r in ranges])
    lo_c, hi_c = (C.c_int * world)(), (C.c_int * world)()
    pkg.check(L.b200_halo_rows(cmin, cmax, world, rank, blocks.count, n, lo_c, hi_c), "b200_halo_rows")
    check("halo_rows C == Python", (list(lo_c), list(hi_c)) == (list(halo[0]), list(halo[1])), (list(lo_c), halo[0]))
    plane = nx * ny
    sent = sum(h - l for d, (l, h) in enumerate(zip(*halo)) if d != rank)
    check("halo is one plane per neighbour", sent == plane * ((rank > 0) + (rank < world - 1)), sent)

    def reset():
        ctx.sync()
        dist.barrier()
        bufs.local[0].upload(x0)
        bufs.local[1].upload(np.full(blocks.padded, -7.0))   # sentinel: rows never sent keep it
        ctx.sync()
        dist.barrier()

    def verify(name, norm, xbuf_ptr, normalised):
        got = pkg.DeviceArray.from_ptr(ctx, xbuf_ptr, blocks.padded, np.float64).download()
        own = got[b0:b1] if normalised else got[b0:b1] / norm
        check(name + " norm", abs(norm - nrm) <= 1e-12 * nrm, f"{norm!r} vs {nrm!r}")
        check(name + " own block", float(np.max(np.abs(own - x[b0:b1]))) <= 1e-12)
        return got

    # "fused" runs twice: with the library's choice of kernel (one chunk per warp at this size) and with the
    # persistent pipelined kernel forced (B200_BCAST_U = 3), which large blocks take by default
    for mode, graph, bcast_u in (("fused", 0, None), ("fused", 4, None), ("fused", 0, 3), ("fused", 4, 3),
                                 ("allgather", 0, None), ("allgather", 4, None)):
        reset()
        ctx.set_option("B200_BCAST_U", bcast_u)
        if bcast_u is not None:
            mode_name = f"{mode}(pipelined)"
        else:
            mode_name = mode
        it = pkg.Iterator(pkg, ctx, comm, sell if mode == "fused" else csr, blocks, rank, world, bufs.ptrs[:2],
                          mode=mode, halo=halo if mode == "fused" else None, graph_steps=graph)
        it.run(7)
        it.run(steps - 7)           # 24 more: graph of 4 replayed six times, or 24 direct steps
        norm = it.norm()
        k, xptr, launches = it.state()
        check(f"{mode_name}/g{graph} steps", k == steps and launches > 0, (k, launches))
        got = verify(f"{mode_name}/g{graph}", norm, xptr, normalised=(mode == "allgather"))
        if mode == "fused":
            # only this rank's block and its neighbours' halo planes were ever written
            need_lo, need_hi = max(b0 - plane, 0), min(b1 + plane, n)
            outside = np.ones(blocks.padded, bool)
            outside[need_lo:need_hi] = False
            check(f"{mode_name}/g{graph} nothing outside block + halo", np.all(got[outside] == -7.0))
            check(f"{mode_name}/g{graph} block + halo complete", not np.any(got[need_lo:need_hi] == -7.0))
        else:
            check(f"{mode_name}/g{graph} whole vector gathered", float(np.max(np.abs(got[:n] - x))) <= 1e-12)
        it.close()
    ctx.set_option("B200_BCAST_U", None)

    # the same fused kernel with the all-reduce + barrier done by the NVSwitch (multimem.red on a multicast
    # block) instead of NCCL: two launches per step, no collective call.  Skipped (recorded) where the box
    # cannot set a multicast object up.
    sup = C.c_int(0)
    pkg.check(L.b200_mcast_supported(ctx.h, C.byref(sup)), "b200_mcast_supported")
    flags = [None] * world
    dist.all_gather_object(flags, sup.value)
    res["mcast"] = "not supported by the device / driver"
    if all(flags):
        try:
            mc = pkg.McastBlock(pkg, ctx, rank, world, all_gather_object=dist.all_gather_object, barrier=dist.barrier)
        except pkg.B200Error as e:
            mc = None
            res["mcast"] = f"set-up refused: {e}"
        if mc is not None:
            res["mcast"] = "ran"
            for graph in (0, 4):
                reset()
                it = pkg.Iterator(pkg, ctx, None, sell, blocks, rank, world, bufs.ptrs[:2], mode="fused_mcast", halo=halo,
                                  graph_steps=graph, mcast=mc)
                it.run(7)
                it.run(steps - 7)
                norm = it.norm()
                k, xptr, launches = it.state()
                check(f"fused_mcast/g{graph} steps", k == steps and launches == 2 * steps, (k, launches))
                verify(f"fused_mcast/g{graph}", norm, xptr, normalised=False)
                it.close()
            ctx.sync()
            dist.barrier()
            mc.close()

    # the sell all-gather formulation, SELL block
    reset()
    it = pkg.Iterator(pkg, ctx, comm, sell, blocks, rank, world, bufs.ptrs[:2], mode="allgather", graph_steps=2)
    it.run(steps)
    verify("allgather-sell/g2", it.norm(), it.state()[1], True)
    it.close()

    # the collective-free ring (flags through peer memory)
    reset()
    bufs.local[1].fill_bytes(0)
    ctx.sync()
    dist.barrier()
    bufs.ring_step = 0
    r = pkg.power_iteration_ring(pkg, ctx, sell, bufs, rank, blocks, steps, halo=halo)
    verify("ring", r.norm, r.x.ptr, False)

    # library collectives by themselves
    t = ctx.array(np.arange(8, dtype=np.float64) + rank)
    comm.allreduce_sum(t.ptr, 8)
    check("allreduce", np.array_equal(t.download(), world * np.arange(8) + sum(range(world))))
    g = ctx.array(np.where(np.arange(4 * world) // 4 == rank, rank + 1.0, 0.0))
    comm.allgather(g.ptr, 4)
    check("allgather", np.array_equal(g.download(), np.repeat(np.arange(world) + 1.0, 4)))
    # ... and for any element type: 12 float32 per rank (48 bytes), the sharded upload of a replicated x
    gb = ctx.array(np.where(np.arange(12 * world) // 12 == rank, rank + 0.5, -1.0).astype(np.float32))
    comm.allgather_bytes(gb.ptr, 48)
    check("allgather_bytes", np.array_equal(gb.download(), np.repeat(np.arange(world) + 0.5, 12).astype(np.float32)))
    comm.check()

    # single-shot, row-sharded: every format's y block equals the oracle's rows
    xs = np.random.default_rng(9).uniform(-1, 1, n)
    y_ref = O.yref(n, rows, cols, vals, xs)
    for dtype, tol in ((np.float64, 1e-12), (np.float32, 1e-5)):
        mats = pkg.build_all(coo, dtype)
        xd = ctx.array(xs.astype(dtype))
        for name, m in mats.items():
            yd = ctx.array(np.full(b1 - b0, np.nan, dtype))
            m.spmv(xd, yd)
            err = float(np.max(np.abs(yd.download() - y_ref[b0:b1])) / np.max(np.abs(y_ref)))
            check(f"shard {name} {np.dtype(dtype).name}", err <= tol, err)

    ctx.sync()
    dist.barrier()
    bufs.close()
    comm.close()
    ctx.close()
    res["ok"] = all(c["ok"] for c in res["checks"].values())
    Path(out).write_text(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if res["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
