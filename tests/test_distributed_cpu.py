"""CPU, world_size 2 over gloo: the host-side logic of the N > 1 paths.

(1) row partition + per-shard format builds: the shards of a row-partitioned matrix, each
    multiplied by the replicated x on its own rank, reassemble to the unsharded result;
(2) the iterated (power-iteration) mode: the same `power_iteration` step logic the GPU path runs,
    here wired to numpy/gloo callables, must match a single-process run to rounding.

The local SpMV in these tests is the CPU oracle -- allowed here because this is tests/; the
product wiring (`gpu_callables`) uses only C-ABI kernels."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def laplace7(nx, ny, nz):
    """Row-sorted triples of the 7-point Laplacian (x fastest): the host twin of
    b200_gen_laplace7_coo, diag 6 / off-diag -1, columns ascending."""
    n = nx * ny * nz
    idx = np.arange(n)
    x, y, z = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    parts = []
    for cond, off in ((z > 0, -nx * ny), (y > 0, -nx), (x > 0, -1), (idx >= 0, 0), (x < nx - 1, 1),
                      (y < ny - 1, nx), (z < nz - 1, nx * ny)):
        r = idx[cond]
        parts.append((r, r + off, np.full(r.size, 6.0 if off == 0 else -1.0)))
    rows = np.concatenate([p[0] for p in parts])
    cols = np.concatenate([p[1] for p in parts])
    vals = np.concatenate([p[2] for p in parts])
    order = np.lexsort((cols, rows))
    return n, rows[order].astype(np.int32), cols[order].astype(np.int32), vals[order]


def numpy_callables(O, ptr, cols, vals, n_local):
    def spmv_local(x_full, seg):
        seg[:n_local] = torch.from_numpy(O.spmv_csr(n_local, ptr, cols, vals, x_full.numpy()))

    def sumsq(seg):
        return (seg[:n_local] ** 2).sum().reshape(1)

    def scale_inv_sqrt(seg, acc):
        seg[:n_local] *= 1.0 / torch.sqrt(acc[0])

    def all_reduce_sum(acc):
        if dist.is_initialized():
            dist.all_reduce(acc)

    def all_gather_inplace(full, seg):
        if dist.is_initialized():
            dist.all_gather_into_tensor(full, seg.clone())

    return dict(spmv_local=spmv_local, sumsq=sumsq, scale_inv_sqrt=scale_inv_sqrt,
                all_reduce_sum=all_reduce_sum, all_gather_inplace=all_gather_inplace)


def run_iteration(rank, world, grid, steps):
    from __graft_entry__ import load_package
    from oracle import binding as O
    pkg = load_package()
    n, rows, cols, vals = laplace7(*grid)
    blocks = pkg.equal_row_blocks(n, world, align=32)
    lo, hi = blocks.bounds(rank)
    sel = (rows >= lo) & (rows < hi)
    ptr, _ = O.build_csr(hi - lo, rows[sel] - lo) if hi > lo else (np.zeros(1, np.int32), 0)
    O.lib().orc_set_threads(1)
    x0 = np.zeros(blocks.padded)
    x0[:n] = np.random.default_rng(1).uniform(0.0, 1.0, n)
    x_cur, x_next = torch.from_numpy(x0.copy()), torch.zeros(blocks.padded, dtype=torch.float64)
    res = pkg.power_iteration(x_cur=x_cur, x_next=x_next, rank=rank, blocks=blocks, steps=steps,
                              **numpy_callables(O, ptr, cols[sel], vals[sel], hi - lo))
    return res.norm, res.x.numpy()[:n].copy()


def _worker(rank, world, port, grid, steps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        norm, x = run_iteration(rank, world, grid, steps)
        np.save(Path(out_dir) / f"x_{rank}.npy", x)
        np.save(Path(out_dir) / f"norm_{rank}.npy", np.array([norm]))
        # (1) single-shot sharded SpMV with replicated x, gathered for checking only
        from __graft_entry__ import load_package
        from oracle import binding as O
        pkg = load_package()
        from conftest import random_sorted_matrix
        n_rows, n_cols = 4096, 4096
        rows, cols, vals = random_sorted_matrix(n_rows, n_cols, 1, 50, 77)
        ptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n_rows))]).astype(np.int32)
        cuts = pkg.partition_rows(ptr, world, align=32)
        r0, r1 = int(cuts[rank]), int(cuts[rank + 1])
        sel = slice(ptr[r0], ptr[r1])
        xv = np.arange(n_cols, dtype=np.float64)
        lptr, _ = O.build_csr(r1 - r0, rows[sel] - r0)
        y_local = torch.from_numpy(O.spmv_csr(r1 - r0, lptr, cols[sel], vals[sel], xv))
        # nnz-balanced blocks have unequal row counts: pad to the largest for the gather
        sizes = [int(cuts[g + 1] - cuts[g]) for g in range(world)]
        padded = torch.zeros(max(sizes), dtype=torch.float64)
        padded[:sizes[rank]] = y_local
        outs = [torch.empty(max(sizes), dtype=torch.float64) for _ in sizes]
        dist.all_gather(outs, padded)
        if rank == 0:
            y = torch.cat([o[:s] for o, s in zip(outs, sizes)]).numpy()
            y_ref = O.yref(n_rows, rows, cols, vals, xv)
            np.save(Path(out_dir) / "shard_err.npy", np.array([O.rel_maxnorm(y, y_ref)]))
    finally:
        dist.destroy_process_group()


def test_power_iteration_world2_matches_single_process(tmp_path):
    grid, steps, world = (12, 10, 9), 25, 2
    norm1, x1 = run_iteration(0, 1, grid, steps)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, grid, steps, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        x = np.load(tmp_path / f"x_{r}.npy")
        norm = float(np.load(tmp_path / f"norm_{r}.npy")[0])
        assert abs(norm - norm1) <= 1e-12 * norm1
        assert np.max(np.abs(x - x1)) <= 1e-12
    assert abs(np.linalg.norm(x1) - 1.0) < 1e-12
    # the estimate approaches the largest eigenvalue of the 7-point Laplacian from below (< 12)
    assert 9.0 < norm1 < 12.0
    assert float(np.load(tmp_path / "shard_err.npy")[0]) <= 1e-12


def _halo_worker(rank, world, port, case, steps, out_dir):
    """Power iteration where every rank only ever RECEIVES the rows halo_rows says it reads: the rest
    of its x buffer stays NaN.  If halo_rows under-estimated a halo, a NaN would enter the SpMV."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from __graft_entry__ import load_package
        from oracle import binding as O
        pkg = load_package()
        O.lib().orc_set_threads(1)
        n, rows, cols, vals, x0 = halo_case(case)
        blocks = pkg.equal_row_blocks(n, world, align=32)
        lo, hi = blocks.bounds(rank)
        sel = (rows >= lo) & (rows < hi)
        ptr, _ = O.build_csr(hi - lo, rows[sel] - lo)
        mine = (int(cols[sel].min()), int(cols[sel].max()))         # what b200_minmax_i32 returns on the GPU
        ranges = [None] * world
        dist.all_gather_object(ranges, mine)
        send_lo, send_hi = pkg.halo_rows(ranges, blocks, rank)
        recv = [pkg.halo_rows(ranges, blocks, d) for d in range(world)]   # what the others send to me
        x = np.full(blocks.padded, np.nan)
        need_lo, need_hi = min(mine[0], lo), max(mine[1] + 1, hi)
        x[need_lo:need_hi] = x0[need_lo:need_hi]                    # own block + halo only
        norm = None
        for _ in range(steps):
            y = O.spmv_csr(hi - lo, ptr, cols[sel], vals[sel], np.where(np.isnan(x), np.nan, x))
            assert not np.isnan(y).any(), "the SpMV read a row nobody sent: halo too small"
            acc = torch.tensor([float((y ** 2).sum())], dtype=torch.float64)
            dist.all_reduce(acc)
            norm = float(acc[0]) ** 0.5
            y = y / norm
            nxt = np.full(blocks.padded, np.nan)
            nxt[lo:hi] = y
            reqs = []
            for d in range(world):                                   # the kernel's peer stores
                if d != rank and send_hi[d] > send_lo[d]:
                    reqs.append(dist.isend(torch.from_numpy(y[send_lo[d]:send_hi[d]].copy()), d))
            for d in range(world):
                r_lo, r_hi = recv[d][0][rank], recv[d][1][rank]
                if d != rank and r_hi > r_lo:
                    buf = torch.empty(r_hi - r_lo, dtype=torch.float64)
                    dist.recv(buf, d)
                    d_lo = blocks.bounds(d)[0]
                    nxt[d_lo + r_lo:d_lo + r_hi] = buf.numpy()
            for q in reqs:
                q.wait()
            x = nxt
        np.save(Path(out_dir) / f"halo_{case}_{rank}.npy", np.concatenate([[norm, lo, hi], x[lo:hi]]))
    finally:
        dist.destroy_process_group()


def halo_case(case):
    if case == "laplace":
        n, rows, cols, vals = laplace7(10, 8, 16)
    else:   # irregular band: every row reaches a different distance to the left and to the right
        rng = np.random.default_rng(33)
        n = 1500
        left, right = rng.integers(0, 140, n), rng.integers(0, 90, n)
        rr, cc = [], []
        for i in range(n):
            c = np.unique(np.clip(np.concatenate([[i], i - rng.integers(0, left[i] + 1, 4),
                                                  i + rng.integers(0, right[i] + 1, 4)]), 0, n - 1))
            rr.append(np.full(c.size, i))
            cc.append(c)
        rows, cols = np.concatenate(rr).astype(np.int32), np.concatenate(cc).astype(np.int32)
        vals = rng.uniform(0.1, 1.0, rows.size)
    x0 = np.random.default_rng(2).uniform(0.0, 1.0, n)
    return n, rows, cols, vals, x0


@pytest.mark.parametrize("case", ["laplace", "irregular"])
def test_halo_limited_exchange_world2(tmp_path, case):
    """World size 2 over gloo: the halo-limited exchange (each rank receives only the rows halo_rows
    assigns to it, everything else in its x buffer is NaN) reproduces the single-process power
    iteration on a stencil and on an irregular band."""
    from oracle import binding as O
    steps, world = 12, 2
    n, rows, cols, vals, x0 = halo_case(case)
    ptr, _ = O.build_csr(n, rows)
    x = x0.copy()
    for _ in range(steps):
        y = O.spmv_csr(n, ptr, cols, vals, x)
        nrm = np.linalg.norm(y)
        x = y / nrm
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_halo_worker, args=(world, port, case, steps, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got = np.load(tmp_path / f"halo_{case}_{r}.npy")
        norm, lo, hi = got[0], int(got[1]), int(got[2])
        assert abs(norm - nrm) <= 1e-12 * nrm
        assert np.max(np.abs(got[3:] - x[lo:hi])) <= 1e-12


def test_equal_row_blocks():
    from __graft_entry__ import load_package
    pkg = load_package()
    b = pkg.equal_row_blocks(64_000_000, 8)
    assert b.count == 8_000_000 and b.padded == 64_000_000 and b.bounds(7) == (56_000_000, 64_000_000)
    b = pkg.equal_row_blocks(1000, 3)
    assert b.count % 32 == 0 and b.count * 3 >= 1000
    assert [b.bounds(r) for r in range(3)] == [(0, 352), (352, 704), (704, 1000)]
    b = pkg.equal_row_blocks(40, 4)       # more ranks than blocks of 32: trailing ranks own nothing
    assert [b.bounds(r) for r in range(4)] == [(0, 32), (32, 40), (40, 40), (40, 40)]


def test_halo_rows_from_column_ranges():
    """halo_rows: the rows of a rank's block that every other rank reads as columns (host logic of
    the halo-limited fused exchange)."""
    from __graft_entry__ import load_package
    pkg = load_package()
    blocks = pkg.equal_row_blocks(1000, 4)            # 256 rows per rank, last block 232
    ranges = [(0, 300), (200, 600), (400, 800), (700, 999)]
    assert pkg.halo_rows(ranges, blocks, 1) == ([0, 0, 144, 0], [45, 256, 256, 0])
    assert pkg.halo_rows(ranges, blocks, 3) == ([0, 0, 0, 0], [0, 0, 33, 232])
    # a rank whose columns span everything gets every block in full; nobody else gets anything
    full = [(0, 999), (300, 400), (600, 700), (800, 900)]
    for r in range(4):
        lo, hi = pkg.halo_rows(full, blocks, r)
        b0, b1 = blocks.bounds(r)
        assert (lo[0], hi[0]) == (0, b1 - b0)
        assert all(hi[d] - lo[d] == 0 for d in range(1, 4) if d != r)
    # 7-point Laplacian blocks: one plane to each neighbour
    nx = ny = 10
    blocks = pkg.equal_row_blocks(nx * ny * 40, 4, align=32)
    ranges = [(max(blocks.bounds(r)[0] - nx * ny, 0), min(blocks.bounds(r)[1] + nx * ny, blocks.n_rows) - 1)
              for r in range(4)]
    lo, hi = pkg.halo_rows(ranges, blocks, 2)
    assert [h - l for l, h in zip(lo, hi)] == [0, 100, blocks.count, 100]


def test_c_abi_halo_rows_equals_the_python_mirror():
    """b200_halo_rows (what the C drivers call, host/src/driver_iterate.c) against halo_rows on random
    column ranges, including ranks that own nothing and ranges that miss a block entirely."""
    import ctypes as C
    import numpy as np
    from __graft_entry__ import load_package
    pkg = load_package()
    L = pkg.lib()
    rng = np.random.default_rng(0)
    for world, n in ((1, 100), (2, 1000), (4, 1000), (8, 40), (8, 123457), (16, 5000)):
        blocks = pkg.equal_row_blocks(n, world)
        for _ in range(20):
            a = rng.integers(0, n, world)
            b = rng.integers(0, n, world)
            ranges = [(int(min(x, y)), int(max(x, y))) for x, y in zip(a, b)]
            cmin = (C.c_int * world)(*[r[0] for r in ranges])
            cmax = (C.c_int * world)(*[r[1] for r in ranges])
            for rank in range(world):
                lo, hi = (C.c_int * world)(), (C.c_int * world)()
                assert L.b200_halo_rows(cmin, cmax, world, rank, blocks.count, n, lo, hi) == 0
                want = pkg.halo_rows(ranges, blocks, rank)
                assert (list(lo), list(hi)) == (list(want[0]), list(want[1])), (world, n, rank, ranges)


def _sharded_x_worker(rank, world, port, out_dir):
    """The sharded upload of a replicated x (bench.py --workload rmat at N > 1, b200_comm_allgather_bytes on the
    GPU): every rank holds only ITS slice of the host's x, the slices are all-gathered into a padded buffer,
    and the row-sharded SpMV with the gathered x must equal the oracle's rows."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from __graft_entry__ import load_package
        from oracle import binding as O
        pkg = load_package()
        O.lib().orc_set_threads(1)
        rng = np.random.default_rng(21)
        n_rows, n_cols = 1111, 1237                                # neither divisible by the world size nor by 4
        lens = rng.integers(1, 9, n_rows)                           # no empty rows: the reference's CSR build (oracle) needs none
        rows = np.repeat(np.arange(n_rows, dtype=np.int32), lens)
        cols = rng.integers(0, n_cols, rows.size).astype(np.int32)
        vals = rng.uniform(-1, 1, rows.size)
        x_host = rng.uniform(0, 1, n_cols).astype(np.float32)      # fp32: 4 entries per 16-byte unit
        per, padded = pkg.x_upload_slices(n_cols, world, 4)
        lo = rank * per
        mine = np.zeros(per, np.float32)
        mine[:max(0, min(per, n_cols - lo))] = x_host[lo:lo + per]                  # what this rank "uploads"
        outs = [torch.empty(per, dtype=torch.float32) for _ in range(world)]
        dist.all_gather(outs, torch.from_numpy(mine))
        x = torch.cat(outs).numpy()
        assert x.size == padded and np.array_equal(x[:n_cols], x_host) and not x[n_cols:].any()
        ptr_host = np.concatenate(([0], np.cumsum(lens))).astype(np.int32)
        cuts = pkg.partition_rows(ptr_host, world, 32)
        r0, r1 = int(cuts[rank]), int(cuts[rank + 1])
        sel = slice(ptr_host[r0], ptr_host[r1])
        lptr, _ = O.build_csr(r1 - r0, rows[sel] - r0)
        y = O.spmv_csr(r1 - r0, lptr, cols[sel], vals[sel], x[:n_cols].astype(np.float64))
        y_ref = O.yref(n_rows, rows, cols, vals, x_host.astype(np.float64))
        np.save(Path(out_dir) / f"sharded_x_{rank}.npy", np.array([O.rel_maxnorm(y, y_ref[r0:r1]) if r1 > r0 else 0.0]))
    finally:
        dist.destroy_process_group()


def test_sharded_x_upload_world2(tmp_path):
    import importlib
    pkg = importlib.import_module("__graft_entry__").load_package()
    for n, world, itemsize in ((1237, 2, 4), (1 << 24, 8, 4), (1001, 3, 8), (16, 8, 8), (5, 4, 4)):
        per, padded = pkg.x_upload_slices(n, world, itemsize)
        assert padded == per * world >= n and (per * itemsize) % 16 == 0
        assert (per - 16 // itemsize) * world < n or per * itemsize == 16           # no more padding than one unit per rank
    assert pkg.x_upload_slices(1 << 24, 8, 4) == (1 << 21, 1 << 24)
    world = 2
    port = 29500 + (os.getpid() % 2000) + 7
    mp.spawn(_sharded_x_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert float(np.load(tmp_path / f"sharded_x_{r}.npy")[0]) <= 1e-12
