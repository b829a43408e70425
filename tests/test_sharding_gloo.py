"""CPU, world_size 2, gloo: the N>1 host logic. k-NN queries shard by contiguous runs of a spatial
(cell-sorted) order with the point set replicated (no collective), every rank returning the rows of
the points it owns plus their caller indices; a Jacobi repel sweep shards the movable points by
contiguous caller-order range and all-gathers the moved positions each iteration. No GPU here, so the per-rank compute
is the CPU oracle standing in for the device kernels; what belongs to the PRODUCT and is exercised through the C ABI of
libwtp_cuda.so in every rank is the partition itself (wtp_shard_begin / wtp_shard_end: the ranges must tile [0, n)
exactly for every n and world, and agree with the host mirror's shard_range) and bench.py's parity sampler
(`sample_rows_brute_force`, the self-check every rank runs under world > 1). The device side of the same paths is
tests/test_gpu_multi.py (2 and 4 GPUs)."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    pkg, oracle = ge.load_package(), ge.load_oracle()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    pts = rng.random((4001, 3))
    # --- k-NN: every rank answers a contiguous run of the cell-sorted order (wtp_shard_owned: caller ids of its
    # rows); scattering every rank's (ids, rows) gives the full table, each row exactly once
    cells = np.floor(pts * 12).astype(np.int64)
    order = np.lexsort((np.arange(len(pts)), cells[:, 0], cells[:, 1], cells[:, 2]))   # row-major cells, stable
    b, e = pkg.shard_range(len(pts), rank, world)
    own = order[b:e]
    mine = oracle.knn(pts, 9, threads=1)[own]
    parts = [None] * world
    dist.all_gather_object(parts, (own, mine))
    full = np.zeros((len(pts), 9), dtype=np.int64)
    hits = np.zeros(len(pts), dtype=np.int64)
    for ids, rows in parts:
        full[ids] = rows
        hits[ids] += 1
    ok_knn = np.array_equal(full, oracle.knn(pts, 9, threads=1)) and (hits == 1).all()
    # --- repel: owned movable range per rank, all-gather of moved positions per iteration
    n_fixed, iters = 401, 4
    h = len(pts) ** (-1 / 3)
    sp, _ = oracle.make_spacing("constant", h)
    f = oracle.make_force("clipped", 0.2)
    kw = dict(alpha_lo=h / 2000, alpha_max=h / 20, tol=0.0, stall_after=0, threads=1)
    snap = pts.copy()
    n_move = len(pts) - n_fixed
    mb, me = pkg.shard_range(n_move, rank, world)
    convs = []
    for _ in range(iters):
        new, conv, _, _ = oracle.repel(snap, n_fixed, sp, f, max_iters=1, **kw)
        owned = torch.from_numpy(new[n_fixed + mb:n_fixed + me].copy())
        sizes = [pkg.shard_range(n_move, r, world) for r in range(world)]
        bufs = [torch.empty((s[1] - s[0], 3), dtype=torch.float64) for s in sizes]
        dist.all_gather(bufs, owned)
        snap[n_fixed:] = torch.cat(bufs).numpy()
        convs.append(conv[0])
    ref, rconv, _, _ = oracle.repel(pts, n_fixed, sp, f, max_iters=iters, **kw)
    ok_repel = np.array_equal(snap, ref)
    # --- the product's partition through the C ABI (no GPU needed): every rank reports its own range, together they tile [0, n)
    lib = pkg._lib.load()
    ok_part = True
    for n in (0, 1, 7, 4001, 10_000_000, (1 << 32) - 17):
        mine = (int(lib.wtp_shard_begin(n, rank, world)), int(lib.wtp_shard_end(n, rank, world)))
        ranges = [None] * world
        dist.all_gather_object(ranges, mine)
        ok_part = ok_part and mine == pkg.shard_range(n, rank, world) and ranges[0][0] == 0 and ranges[-1][1] == n \
            and all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1)) and all(a <= b for a, b in ranges)
    # --- bench.py's parity sampler on this rank's rows (brute force in numpy against the table the ranks assembled)
    import bench
    qi = own[:: max(len(own) // 16, 1)][:16]
    ok_sampler = np.array_equal(bench.sample_rows_brute_force(pts, qi, 9), full[qi])
    ret[rank] = (ok_knn, ok_repel, ok_part, ok_sampler)
    dist.destroy_process_group()


def test_sharded_knn_and_repel_world2():
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, 29611, ret), nprocs=world, join=True)
    assert all(ret[r] == (True, True, True, True) for r in range(world)), dict(ret)
