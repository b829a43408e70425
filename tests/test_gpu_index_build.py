"""GPU (-m gpu): the two index builds of grid.cu — counting sort (default) and the stable radix sort (WTP_RADIX_BUILD=1)
— must produce the SAME sorted order: cells ascending,
ascending caller index inside a cell. wtp_shard_owned on an unsharded context returns that order (row t of the
compact table = sorted position t), so the test reads it back after a k-NN call under either build.

Covers the three placement paths: cells of a few points (ranked by counting), heavy cells (one CTA sorts the cell in
shared memory), and cells above the shared-memory size (the CTA sorts them in global memory)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _clouds():
    rng = np.random.default_rng(77)
    out = {}
    out["uniform3d_f32"] = rng.random((200_000, 3), dtype=np.float32)
    out["uniform2d_f64"] = rng.random((50_000, 2))
    # a ball of 3000 points inside one cell + 100 copies of one point: cells far above the ranking limit
    base = rng.random((30_000, 3), dtype=np.float32)
    ball = (np.float32(0.5) + rng.random((3000, 3), dtype=np.float32) * np.float32(1e-4))
    same = np.repeat(np.array([[0.25, 0.75, 0.5]], dtype=np.float32), 100, axis=0)
    heavy = np.concatenate([base, ball, same])
    out["heavy_cells_f32"] = heavy[rng.permutation(len(heavy))]
    # cells heavier than the shared-memory sort takes (8192): sorted in global memory, 9000 and 20000 points (not powers of two)
    blob = (0.5 + rng.random((9000, 3)) * 1e-7)
    over = np.concatenate([rng.random((20_000, 3)), blob])
    out["over_limit_f64"] = over[rng.permutation(len(over))]
    blob2 = (np.float32(0.3) + rng.random((20_000, 3), dtype=np.float32) * np.float32(1e-6))
    over2 = np.concatenate([rng.random((30_000, 3), dtype=np.float32), blob2, ball])
    out["far_over_limit_f32"] = over2[rng.permutation(len(over2))]
    return out


@pytest.mark.parametrize("name", ["uniform3d_f32", "uniform2d_f64", "heavy_cells_f32", "over_limit_f64", "far_over_limit_f32"])
def test_counting_build_equals_radix_build(ctx, name):
    pts = _clouds()[name]
    k = 5
    old = os.environ.pop("WTP_RADIX_BUILD", None)
    try:
        idx_c = ctx.knn(pts, k)
        order_c = ctx.owned().copy()
        os.environ["WTP_RADIX_BUILD"] = "1"
        idx_r = ctx.knn(pts, k)
        order_r = ctx.owned().copy()
    finally:
        os.environ.pop("WTP_RADIX_BUILD", None)
        if old is not None:
            os.environ["WTP_RADIX_BUILD"] = old
    n = len(pts)
    assert len(order_c) == n and np.array_equal(np.sort(order_c), np.arange(1, n + 1))
    assert np.array_equal(order_c, order_r)
    assert np.array_equal(idx_c, idx_r)


def test_counting_build_radius_rows(ctx, oracle):
    """Radius rows come out in the sorted order's scan order merged by caller index: the CSR must equal the oracle's
    under the counting build too (heavy cells included)."""
    pts = _clouds()["heavy_cells_f32"]
    off, ind = ctx.radius(pts, 0.03)
    ro, ri = oracle.radius(pts, 0.03)
    assert np.array_equal(off, ro) and np.array_equal(ind, ri)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_int32_device_table_equals_int64_table(ctx, dt):
    """wtp_knn_dev_i32_*: the same rows as 4-byte indices (device pointers in and out), distances included, tiled pass
    and general kernel (k = 40) alike; a sharded context writes the same compact table."""
    import torch
    rng = np.random.default_rng(9)
    pts = rng.random((120_000, 3)).astype(dt)
    dev = torch.device("cuda", 0)
    d = torch.from_numpy(pts).to(dev)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    try:
        for k in (21, 40):
            i64 = torch.empty((len(pts), k), dtype=torch.int64, device=dev)
            i32 = torch.full((len(pts), k), -1, dtype=torch.int32, device=dev)
            dist = torch.empty((len(pts), k), dtype=torch.float32 if dt == np.float32 else torch.float64, device=dev)
            dist32 = torch.empty_like(dist)
            ctx.knn_dev(d.data_ptr(), len(pts), 3, k, dt, i64.data_ptr(), dist.data_ptr())
            ctx.knn_dev(d.data_ptr(), len(pts), 3, k, dt, i32.data_ptr(), dist32.data_ptr(), idx32=True)
            torch.cuda.synchronize()
            assert torch.equal(i64, i32.to(torch.int64)) and torch.equal(dist, dist32)
            i32n = torch.full((len(pts), k), -1, dtype=torch.int32, device=dev)
            ctx.knn_dev(d.data_ptr(), len(pts), 3, k, dt, i32n.data_ptr(), 0, idx32=True)     # indices only (the row-per-store path)
            torch.cuda.synchronize()
            assert torch.equal(i32, i32n)
    finally:
        ctx.set_stream(None)
