"""GPU (-m gpu): the CUDA path through the C ABI against the CPU oracle on identical inputs.
Bit-exact for neighbour lists (indices AND distances); repel positions within the north-star
tolerance (1e-6*spacing Float64, 1e-3*spacing Float32 after 10 iterations) — in practice to
rounding (1e-12 / 2e-5 of a spacing): the sweep uses the same neighbours in the same order and
differs only in how the force terms are rounded (fused multiply-adds, one division per neighbour)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-6, np.float32: 1e-3}


def rows_of(off, ind):
    return [ind[off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]


# ------------------------------------------------------------------ k-NN
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("D", [2, 3])
@pytest.mark.parametrize("N,k", [(50, 5), (1000, 21), (30000, 21), (5000, 40), (3000, 100), (2500, 200), (22, 21), (2, 1)])
def test_knn_bit_exact(ctx, oracle, dt, D, N, k):
    pts = np.random.default_rng(N + k + D).random((N, D)).astype(dt)
    a, ad = ctx.knn(pts, k, dists=True)
    b, bd = oracle.knn(pts, k, dists=True)
    assert np.array_equal(a, b) and np.array_equal(ad, bd)
    assert not (a == np.arange(1, N + 1)[:, None]).any()            # self excluded (test/topology.jl:40)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_knn_clustered_graded_duplicates(ctx, oracle, dt):
    rng = np.random.default_rng(1)
    clustered = np.concatenate([rng.normal(0, 0.01, (5000, 3)), rng.random((5000, 3)) * 10]).astype(dt)
    assert np.array_equal(ctx.knn(clustered, 21), oracle.knn(clustered, 21))
    assert ctx.timing()["n_ring_expanded"] > 0                       # the ring expansion path ran
    graded = (rng.random((20000, 3)) ** 3).astype(dt)                # density varies by orders of magnitude
    assert np.array_equal(ctx.knn(graded, 21), oracle.knn(graded, 21))
    dup = np.repeat(rng.random((500, 3)), 4, axis=0).astype(dt)      # ties: order by index
    assert np.array_equal(ctx.knn(dup, 9), oracle.knn(dup, 9))
    flat = rng.random((4000, 3)).astype(dt); flat[:, 2] = 0.5        # degenerate extent
    assert np.array_equal(ctx.knn(flat, 12), oracle.knn(flat, 12))
    far = (rng.random((3000, 3)) + 1.0e4).astype(dt)                 # large offset: few mantissa bits left in f32
    assert np.array_equal(ctx.knn(far, 10), oracle.knn(far, 10))


def test_knn_real_geometry(ctx, oracle, stl_points):
    for name, pts in stl_points.items():                             # STL face centres
        a, ad = ctx.knn(pts, 21, dists=True)
        b, bd = oracle.knn(pts, 21, dists=True)
        assert np.array_equal(a, b) and np.array_equal(ad, bd), name


def test_search_including_self_and_known_answers(ctx, oracle, known):
    g = known["circle_k3"]                                           # test/neighbors.jl:36-56
    idx, dist = ctx.knn(np.array(g["points"]), 3, include_self=True, dists=True)
    assert (idx[:, 0] == np.arange(1, 21)).all()
    assert [sorted(r) for r in idx.tolist()] == g["sets"]
    assert (dist[:, 0] == 0).all() and (np.diff(dist, axis=1) >= 0).all()
    pts = np.random.default_rng(2).random((5, 3))
    assert ctx.knn(pts, 5, include_self=True).shape == (5, 5)        # k == N (test/neighbors.jl:158-166)
    assert (ctx.knn(pts, 1, include_self=True)[:, 0] == np.arange(1, 6)).all()


def test_knn_errors(ctx, pkg):
    pts = np.random.default_rng(3).random((5, 3))
    with pytest.raises(pkg.WtpArgumentError):                        # k + 1 > N
        ctx.knn(pts, 5)
    with pytest.raises(pkg.WtpArgumentError):
        ctx.knn(np.random.rand(500, 3), 300)                         # above WTP_MAX_K
    with pytest.raises(pkg.WtpArgumentError):
        ctx.knn(np.random.rand(10, 4), 2)


def test_knn_host_pipeline_variants(ctx, oracle, monkeypatch):
    """The host entry point's ways back to the caller's int64 table: 3-byte packed indices (every index below 2^24, the
    default here), plain 4-byte indices (WTP_NO_PACK24: what larger point sets use), other staging geometries."""
    pts = np.random.default_rng(77).random((300_000, 3)).astype(np.float32)      # 6.3 M entries: the chunked pipeline, not the small-result copy
    ref = oracle.knn(pts, 21)
    assert np.array_equal(ctx.knn(pts, 21), ref)
    monkeypatch.setenv("WTP_NO_PACK24", "1")
    assert np.array_equal(ctx.knn(pts, 21), ref)
    monkeypatch.delenv("WTP_NO_PACK24")
    for mb, slots in (("1", "3"), ("16", "2")):
        monkeypatch.setenv("WTP_STAGE_MB", mb); monkeypatch.setenv("WTP_STAGE_SLOTS", slots)
        assert np.array_equal(ctx.knn(pts, 21), ref)
    monkeypatch.delenv("WTP_STAGE_MB"); monkeypatch.delenv("WTP_STAGE_SLOTS")
    off, ind = ctx.radius(pts[:200_000], 0.03)                                   # > 4 M entries: the CSR indices take the same path
    roff, rind = oracle.radius(pts[:200_000], 0.03)
    assert len(ind) > (4 << 20) and np.array_equal(off, roff) and np.array_equal(ind, rind)


def test_knn_cell_occupancy_invariance(ctx, oracle):
    pts = np.random.default_rng(4).random((20000, 3)).astype(np.float32)
    ref = oracle.knn(pts, 21)
    for m in (1.0, 3.0, 20.0, 100.0):                                # results must not depend on the grid
        ctx.set_cell_occupancy(m)
        assert np.array_equal(ctx.knn(pts, 21), ref), m
    ctx.set_cell_occupancy(0.0)


# ----------------------------------------------------- sharded k-NN, one rank at a time
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("D", [2, 3])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_knn_shards_windowed_index(pkg, oracle, dt, D, world):
    """Every rank of a sharded context (shard-only: no communicator is needed for k-NN) indexes only its window
    of the grid; the rows it answers are the oracle's, and the ranks' runs tile the sorted order exactly."""
    N = 60001
    pts = np.random.default_rng(world * 10 + D).random((N, D)).astype(dt)
    ref, ref_d = oracle.knn(pts, 21, dists=True)
    seen = np.zeros(N, dtype=np.int32)
    for rank in range(world):
        c = pkg.Context(0)
        c.comm_init(rank, world, None)
        idx = np.zeros((N, 21), dtype=np.int64)
        dist = np.zeros((N, 21), dtype=dt) if world == 3 else None          # searchdists through the sharded host path too
        c.knn(pts, 21, out_idx=idx, out_dist=dist)
        own = c.owned() - 1
        b, e = c.shard(N)
        t = c.timing()
        assert len(own) == e - b and np.array_equal(idx[own], ref[own])
        if dist is not None:
            assert np.array_equal(dist[own], ref_d[own]) and (dist[np.setdiff1d(np.arange(N), own)] == 0).all()
        rest = np.ones(N, dtype=bool); rest[own] = False
        assert (idx[rest] == 0).all()
        assert 0 < t["n_window_points"] < N and t["n_window_missed"] == 0      # the window, not the whole set, was sorted
        seen[own] += 1
        c.close()
    assert (seen == 1).all()


def test_knn_shards_window_fallback(pkg, oracle):
    """A strongly graded cloud: searches leave the window, the call is repeated on the whole index (same rows),
    and the context stops windowing."""
    rng = np.random.default_rng(3)
    pts = (rng.random((40000, 3)) ** 4).astype(np.float32)
    ref = oracle.knn(pts, 21)
    seen = np.zeros(len(pts), dtype=np.int32)
    missed = 0
    for rank in range(4):
        c = pkg.Context(0)
        c.comm_init(rank, 4, None)
        idx = np.zeros((len(pts), 21), dtype=np.int64)
        c.knn(pts, 21, out_idx=idx)
        own = c.owned() - 1
        assert np.array_equal(idx[own], ref[own])
        t = c.timing()
        missed += t["n_window_missed"]
        if t["n_window_missed"] > 0:                                    # second call: no window attempt any more
            idx2 = np.zeros_like(idx)
            c.knn(pts, 21, out_idx=idx2)
            assert np.array_equal(idx2, idx) and c.timing()["n_window_points"] == 0
        seen[own] += 1
        c.close()
    assert (seen == 1).all() and missed > 0


def test_shard_only_context_refuses_repel(pkg):
    c = pkg.Context(0)
    c.comm_init(0, 2, None)
    pts = np.random.default_rng(0).random((500, 3))
    sp, _ = c.make_spacing("constant", a=0.1)
    with pytest.raises(pkg.WtpError):
        c.repel(pts, 50, sp, c.make_force("clipped", 0.2), max_iters=2, alpha_lo=1e-5, alpha_max=1e-3)
    c.close()


# ---------------------------------------------------------------- radius
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("D,r", [(2, 0.03), (3, 0.08), (2, 0.004), (3, 0.5)])
def test_radius_csr_bit_exact(ctx, oracle, dt, D, r):
    n = 20000 if r < 0.4 else 1500                                   # r = 0.5: rows of hundreds (rank path)
    pts = np.random.default_rng(D).random((n, D)).astype(dt)
    off, ind = ctx.radius(pts, r)
    roff, rind = oracle.radius(pts, r)
    assert np.array_equal(off, roff) and np.array_equal(ind, rind)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("D", [2, 3])
def test_radius_tiled_and_leftover_rows(ctx, oracle, dt, D):
    """Rows around the capacity of the tiled fill's per-thread lists (about 100 hits): short rows are merged in the
    tiled pass, longer ones go to the general kernel; graded density mixes both in one call. Lattice points give
    exact distance ties at the radius (inclusive) and rows whose runs interleave."""
    rng = np.random.default_rng(11 + D)
    n = 40000
    graded = (rng.random((n, D)) ** 2).astype(dt)                        # density varies ~100x across the domain
    r = (60.0 / n) ** (1.0 / D) * (0.55 if D == 2 else 0.6)
    off, ind = ctx.radius(graded, r)
    roff, rind = oracle.radius(graded, r)
    assert np.array_equal(off, roff) and np.array_equal(ind, rind)
    rows = np.diff(off)
    assert rows.max() > 130 and np.median(rows) < 100                    # both sides of the list capacity
    m = 60 if D == 2 else 16
    axes = [np.arange(m, dtype=np.float64) / 8.0] * D                    # spacing 0.125: exactly representable
    lattice = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, D)
    lattice = lattice[rng.permutation(len(lattice))].astype(dt)
    off, ind = ctx.radius(lattice, 0.25)                                 # neighbours at exactly r are hits (d2 <= r2)
    roff, rind = oracle.radius(lattice, 0.25)
    assert np.array_equal(off, roff) and np.array_equal(ind, rind)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("D", [2, 3])
def test_radius_shards_one_rank_at_a_time(pkg, oracle, dt, D):
    """Sharded radius (shard-only contexts, one rank at a time on one GPU): rank r answers the caller range
    [shard_begin, shard_end) with the tiled passes (the range is a keep filter on every tile) and returns that part of
    the CSR, offsets 0-based within the shard; together the shards are the oracle's CSR."""
    rng = np.random.default_rng(300 + D)
    n = 50003
    pts = (rng.random((n, D)) ** 1.5).astype(dt)                        # mildly graded: short and long rows
    r = (40.0 / n) ** (1.0 / D) * 0.6
    roff, rind = oracle.radius(pts, r)
    world = 3
    for rank in range(world):
        c = pkg.Context(0)
        c.comm_init(rank, world, None)
        b, e = c.shard(n)
        off, ind = c.radius(pts, r)
        assert len(off) == e - b + 1 and np.array_equal(off, roff[b:e + 1] - roff[b]) and np.array_equal(ind, rind[roff[b]:roff[e]])
        c.close()


def test_radius_known_answer_and_edges(ctx, oracle, known):
    g = known["radius_grid5x5"]                                      # test/topology.jl:46-52
    for dt in (np.float64, np.float32):
        off, ind = ctx.radius(np.array(g["points"], dtype=dt), g["radius"])
        assert rows_of(off, ind) == g["rows"]
    pts = np.random.default_rng(5).random((1000, 3))
    off, ind = ctx.radius(pts, 1e-9)                                 # nobody in range: empty CSR
    assert off[-1] == 0 and ind.size == 0
    dup = np.repeat(pts[:100], 2, axis=0)
    off, ind = ctx.radius(dup, 0.0)                                  # r = 0: only the coincident twin, self removed by index
    roff, rind = oracle.radius(dup, 0.0)
    assert np.array_equal(off, roff) and np.array_equal(ind, rind) and (np.diff(off) == 1).all()


# ------------------------------------------------------- forces, spacings
def test_compute_force_known_answers(ctx, known):
    g = known["compute_force"]                                       # test/repel.jl:117-170
    for kind, key, u0 in (("inverse", "inverse", 1.0), ("equilibrium", "equilibrium", 1.0), ("clipped", "clipped_u0_1", 1.0),
                          ("clipped", "clipped_u0_0.8", 0.8), ("strong", "strong_gamma3", 1.0)):
        got = ctx.force_eval(ctx.make_force(kind, g["beta"], u0, 3.0), np.array(g["u"]))
        np.testing.assert_allclose(got, g[key], rtol=1e-14, atol=0)
        got32 = ctx.force_eval(ctx.make_force(kind, g["beta"], u0, 3.0), np.array(g["u"], dtype=np.float32))
        assert got32.dtype == np.float32
        np.testing.assert_allclose(got32, g[key], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("D", [2, 3])
def test_spacings_match_oracle(ctx, oracle, dt, D):
    rng = np.random.default_rng(6)
    bnd = rng.random((3000, D)); bnd[:, 0] = 0
    q = (rng.random((20000, D)) * 3 - 1).astype(dt)                  # queries also far outside the boundary set's box
    for kind, a, b, c in (("constant", 0.1, 0, 0), ("loglike", 0.1, 1.5, 0), ("boundary_layer", 0.01, 0.04, 0.2)):
        sp, k1 = ctx.make_spacing(kind, a, b, c, bnd.astype(dt) if kind != "constant" else None)
        osp, k2 = oracle.make_spacing(kind, a, b, c, bnd.astype(dt) if kind != "constant" else None)
        x, y = ctx.spacing_eval(sp, q), oracle.spacing_eval(osp, q)
        if kind == "boundary_layer":                                 # exp(): CUDA and glibc differ in the last ulp
            np.testing.assert_allclose(x, y, rtol=4e-7 if dt == np.float32 else 1e-15)
        else:
            assert np.array_equal(x, y), kind


# ----------------------------------------------------------------- repel
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("D", [2, 3])
@pytest.mark.parametrize("skind", ["constant", "boundary_layer", "loglike"])
def test_repel_10_iterations_within_tolerance(ctx, oracle, dt, D, skind):
    rng = np.random.default_rng(7)
    N, nf = 6000, 800
    snap = rng.random((N, D)).astype(dt)
    h = N ** (-1.0 / D)
    args = {"constant": ("constant", h, 0, 0, None), "boundary_layer": ("boundary_layer", 0.7 * h, 1.3 * h, 0.2, snap[:nf]),
            "loglike": ("loglike", 2 * h, 1.5, 0, snap[:nf])}[skind]
    sp, k1 = ctx.make_spacing(*args)
    osp, k2 = oracle.make_spacing(*args)
    smin = float(oracle.spacing_eval(osp, snap).min())
    kw = dict(k=21, max_iters=10, tol=0.0, stall_after=0, alpha_lo=smin / 2000, alpha_max=smin / 20, trace=True)
    out, conv, res, tr = ctx.repel(snap, nf, sp, ctx.make_force("clipped", 0.2), **kw)
    oout, oconv, ores, otr = oracle.repel(snap, nf, osp, oracle.make_force("clipped", 0.2), **kw)
    assert res["iters"] == ores["iters"] == 10 and res["stop_reason"] == "max_iters"
    assert np.array_equal(out[:nf], snap[:nf])                       # the wall is untouched
    assert np.abs(out - oout).max() <= TOL[dt] * smin
    np.testing.assert_allclose(conv, oconv, rtol=1e-5 if dt == np.float32 else 1e-12)
    assert [(t["idx_a"], t["idx_b"]) for t in tr] == [(t["idx_a"], t["idx_b"]) for t in otr]
    # the same neighbours in the same order; the force terms are accumulated with fused multiply-adds and one division
    # per neighbour on the device, so positions agree to rounding, far inside the tolerance
    assert np.abs(out - oout).max() <= (1e-12 if dt == np.float64 else 2e-5) * smin


@pytest.mark.parametrize("kind", ["inverse", "equilibrium", "clipped", "strong"])
def test_repel_force_models(ctx, oracle, kind):
    rng = np.random.default_rng(8)
    snap = rng.random((3000, 3))
    h = 3000 ** (-1 / 3)
    sp, _ = ctx.make_spacing("constant", h)
    osp, _ = oracle.make_spacing("constant", h)
    kw = dict(max_iters=5, tol=0.0, stall_after=0, alpha_lo=h / 2000, alpha_max=h / 20)
    out = ctx.repel(snap, 300, sp, ctx.make_force(kind, 0.2, 1.0, 3.0), **kw)[0]
    oout = oracle.repel(snap, 300, osp, oracle.make_force(kind, 0.2, 1.0, 3.0), **kw)[0]
    assert np.abs(out - oout).max() <= 1e-6 * h


@pytest.mark.parametrize("kw", [dict(stall_after=1, tol=1e-12, max_iters=600), dict(cv_target=10.0, tol=1e-12, max_iters=50, stall_after=0),
                                dict(tol=1e6, max_iters=50, stall_after=0), dict(rebuild_every=3, max_iters=12, tol=0.0, stall_after=0),
                                dict(k=40, max_iters=4, tol=0.0, stall_after=0), dict(k=5000, max_iters=2, tol=0.0, stall_after=0)])
def test_repel_stop_logic_and_options(ctx, oracle, pkg, kw):
    rng = np.random.default_rng(9)
    n = 100 if kw.get("k", 21) > 128 else 4000                       # k > N: kk = min(k, N) (src/repel.jl:208)
    snap = rng.random((n, 3))
    h = n ** (-1 / 3)
    sp, _ = ctx.make_spacing("constant", h)
    osp, _ = oracle.make_spacing("constant", h)
    a = dict(alpha_lo=h / 2000, alpha_max=h / 20)
    out, conv, res, _ = ctx.repel(snap, n // 8, sp, ctx.make_force("clipped", 0.2), **a, **kw)
    oout, oconv, ores, _ = oracle.repel(snap, n // 8, osp, oracle.make_force("clipped", 0.2), **a, **kw)
    assert (res["iters"], res["stop_reason"]) == (ores["iters"], ores["stop_reason"])
    assert np.abs(out - oout).max() <= 1e-6 * h
    if "cv_target" in kw:
        assert res["iters"] == 1 and np.array_equal(out, snap)       # pre-sweep configuration (test/repel.jl:281-289)


def test_repel_rejects_what_cannot_cross_the_abi(ctx, pkg):
    snap = np.random.default_rng(10).random((500, 3))
    sp, _ = ctx.make_spacing("constant", 0.1)
    f = ctx.make_force("clipped", 0.2)
    with pytest.raises(pkg.WtpArgumentError):
        ctx.repel(snap, 50, sp, f, rebuild_every=0, alpha_lo=1e-4, alpha_max=1e-2)
    with pytest.raises(pkg.WtpArgumentError):
        ctx.repel(snap, 50, sp, f, kick_after=-1, alpha_lo=1e-4, alpha_max=1e-2)
    bad = pkg._lib.Force(9, 0.2, 1.0, 3.0)
    with pytest.raises(pkg.WtpError):
        ctx.repel(snap, 50, sp, bad, alpha_lo=1e-4, alpha_max=1e-2)


def test_converged_metrics_agree(ctx, oracle):
    """North star: after full convergence the metrics agree within 1 %."""
    rng = np.random.default_rng(11)
    snap = rng.random((3000, 3))
    h = 3000 ** (-1 / 3)
    sp, _ = ctx.make_spacing("constant", h)
    osp, _ = oracle.make_spacing("constant", h)
    kw = dict(max_iters=400, tol=1e-12, stall_after=1, alpha_lo=h / 2000, alpha_max=h / 20)
    out, _, res, _ = ctx.repel(snap, 400, sp, ctx.make_force("clipped", 0.2), **kw)
    oout, _, ores, _ = oracle.repel(snap, 400, osp, oracle.make_force("clipped", 0.2), **kw)
    assert res["stop_reason"] == ores["stop_reason"] and abs(res["iters"] - ores["iters"]) <= 1
    m, om = ctx.metrics(out, 20), oracle.metrics(oout, 20)
    for key in om:
        assert abs(m[key] - om[key]) <= 0.01 * abs(om[key]), key
    m2 = oracle.metrics(out, 20)                                     # device metrics kernel vs oracle on the same cloud
    for key in m2:
        assert abs(m[key] - m2[key]) <= 1e-9 * abs(m2[key]), key


# -------------------------------------------------- reference-facing layer
def test_set_topology_api(ctx, pkg, oracle):
    rng = np.random.default_rng(12)
    pts = rng.random((20, 3))
    cloud = pkg.PointCloud(pkg.PointBoundary(pts))
    cloud = pkg.set_topology(cloud, pkg.KNNTopology, 5, ctx=ctx)     # test/topology.jl:12-41
    assert pkg.hastopology(cloud) and isinstance(pkg.topology(cloud), pkg.KNNTopology) and pkg.topology(cloud).k == 5
    nb = pkg.neighbors(cloud)
    assert len(nb) == 20 and all(len(r) == 5 for r in nb)
    assert all(i not in pkg.neighbors(cloud, i) for i in range(1, 21))
    pkg.rebuild_topology_(cloud, ctx=ctx)                            # test/topology.jl:68-84
    assert pkg.topology(cloud).k == 5 and len(pkg.neighbors(cloud, 1)) == 5
    grid = np.array([[i * 0.1, j * 0.1] for i in range(5) for j in range(5)])
    c2 = pkg.set_topology(pkg.PointCloud(grid), pkg.RadiusTopology, 0.15, ctx=ctx)   # test/topology.jl:43-66
    assert pkg.topology(c2).radius == 0.15 and all(i not in pkg.neighbors(c2, i) for i in range(1, 26))
    c3 = pkg.set_topology(pkg.PointCloud(grid), pkg.RadiusTopology, lambda p: 0.15, ctx=ctx)   # radius as f(points), src/topology.jl:100
    assert [r.tolist() for r in pkg.neighbors(c3)] == [r.tolist() for r in pkg.neighbors(c2)]
    vol = pkg.set_topology(pkg.PointVolume(pts), pkg.KNNTopology, 3, ctx=ctx)        # volume-level (test/topology.jl:165+)
    assert np.array_equal(pkg.neighbors(vol).table, oracle.knn(pts, 3))


def test_repel_api(ctx, pkg, oracle):
    rng = np.random.default_rng(13)
    bnd, vol = rng.random((300, 3)).astype(np.float32), rng.random((2000, 3)).astype(np.float32)
    cloud = pkg.PointCloud(bnd, vol)
    sp = pkg.ConstantSpacing(np.float32(0.08))
    conv, trace = [], []
    out = pkg.repel(cloud, sp, beta=np.float32(0.2), max_iters=3, convergence=conv, trace=trace, isinside=False, ctx=ctx)   # test/float32_pipeline.jl:44
    assert pkg.points(out).dtype == np.float32 and len(conv) == 3 and len(trace) == 3
    assert isinstance(pkg.topology(out), pkg.NoTopology) and len(out) == len(cloud)
    assert all(np.isfinite(conv)) and all(c >= 0 for c in conv)
    assert trace[0]["idx_a"] < trace[0]["idx_b"]
    c2 = pkg.repel(cloud, sp, max_iters=100, tol=1.0e6, isinside=False, ctx=ctx)      # test/repel.jl:8-11
    assert c2.repel_result["iters"] == 1
    with pytest.raises(pkg.WtpArgumentError):                         # the survivor filter is on by default (src/repel.jl:90) and needs
        pkg.repel(cloud, sp, max_iters=1, ctx=ctx)                    # the boundary elements' normals and areas in 3-D


# ------------------------------------------------------- mesh wall rule (R6)
def _mesh_queries(mesh, rng, n, dt):
    """Points inside, outside, far away, hugging the surface, and exactly on vertices / edge midpoints."""
    lo, hi = mesh.bbox_min.astype(np.float64), mesh.bbox_max.astype(np.float64)
    ext = hi - lo
    tri = mesh.triangles.reshape(-1, 3, 3).astype(np.float64)
    box = lo + rng.random((n, 3)) * ext
    wide = lo - ext + rng.random((n // 2, 3)) * 3 * ext
    pick = rng.integers(0, len(tri), n)
    bary = rng.dirichlet([1, 1, 1], n)
    on = (tri[pick] * bary[:, :, None]).sum(1)
    near = on + rng.normal(0, 1e-3, (n, 3)) * np.linalg.norm(ext)
    verts = tri[pick[:200], 0]
    mids = 0.5 * (tri[pick[:200], 0] + tri[pick[:200], 1])
    return np.ascontiguousarray(np.concatenate([box, wide, near, on, verts, mids]).astype(dt))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("shape", ["cube", "cuboid", "sphere", "torus"])
def test_mesh_isinside_project_bit_exact(ctx, oracle, pkg, dt, shape):
    mesh = {"cube": lambda: pkg.unit_cube_mesh(dt), "cuboid": lambda: pkg.cuboid_mesh(20.0, 7.0, 3.0, dt),
            "sphere": lambda: pkg.icosphere_mesh(3, 2.5, (1.0, -2.0, 0.5), dt), "torus": lambda: pkg.torus_mesh(dtype=dt)}[shape]()
    q = _mesh_queries(mesh, np.random.default_rng(len(shape)), 3000, dt)
    assert np.array_equal(ctx.mesh_isinside(mesh, q), oracle.mesh_isinside(mesh, q))
    p, tri = ctx.mesh_project(mesh, q)
    op, otri = oracle.mesh_project(mesh, q)
    assert np.array_equal(tri, otri) and np.array_equal(p, op)      # canonical (d2, triangle index) nearest triangle, same arithmetic


def test_mesh_isinside_known_answers(ctx, pkg):
    cube = pkg.unit_cube_mesh()                                      # test/octree_isinside.jl:7-11, 57-63, 106-137
    assert pkg.isinside(np.array([[0.5, 0.5, 0.5], [-0.5, 0.5, 0.5], [0.3, 0.3, 0.3]]), cube, ctx=ctx).tolist() == [True, False, True]
    d = 1.0e-3
    assert pkg.isinside(np.array([[0.5, 0.5, d], [0.5, 0.5, -d], [d, 0.5, d], [-d, 0.5, -d]]), cube, ctx=ctx).tolist() == [True, False, True, False]
    box = pkg.cuboid_mesh(20.0, 7.0, 3.0)                            # :66-103
    assert pkg.isinside(np.array([[5, 3.5, 1.5], [5, 3.5, 10.0], [25, 3.5, 1.5]]), box, ctx=ctx).tolist() == [True, False, False]
    assert pkg.isinside(np.array([0.5, 0.5, 0.5]), cube, ctx=ctx) is True


def _wall_problem(pkg, rng, dt, n_vol=6000, sub=3):
    sph = pkg.icosphere_mesh(sub, dtype=dt)
    bnd = sph.triangles.reshape(-1, 3, 3).astype(np.float64).mean(axis=1)
    vol = rng.normal(size=(n_vol, 3))
    vol *= (0.972 * rng.random((n_vol, 1)) ** (1 / 3)) / np.linalg.norm(vol, axis=1, keepdims=True)   # some start right under the wall
    snap = np.ascontiguousarray(np.concatenate([bnd, vol]).astype(dt))
    return sph, snap, np.arange(len(snap)) < len(bnd)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("skind", ["constant", "boundary_layer"])
def test_repel_mesh_wall_10_iterations(ctx, oracle, pkg, dt, skind):
    """repel(cloud, spacing, octree): n_fixed = 0, boundary points re-projected, escapees reverted and flagged."""
    sph, snap, is_bnd = _wall_problem(pkg, np.random.default_rng(21), dt)
    h = 0.085
    bset = np.ascontiguousarray(snap[is_bnd])
    args = ("constant", h) if skind == "constant" else ("boundary_layer", 0.6 * h, 1.2 * h, 0.4, bset)
    sp, keep = ctx.make_spacing(*args)
    osp, okeep = oracle.make_spacing(*args)
    kw = dict(max_iters=10, tol=0.0, stall_after=0, alpha_lo=h / 2000, alpha_max=h / 2)   # large steps: some points hit the wall
    out, conv, res, _ = ctx.repel(snap, 0, sp, ctx.make_force("clipped", dt(0.2)), mesh=sph, is_bnd=is_bnd, **kw)
    wall = ctx.last_wall
    oout, oconv, ores, _ = oracle.repel(snap, 0, osp, oracle.make_force("clipped", dt(0.2)), mesh=sph, is_bnd=is_bnd, **kw)
    owall = oracle.repel.last_wall
    assert res["iters"] == ores["iters"] == 10
    assert np.array_equal(wall["escaped"], owall["escaped"]) and wall["escaped"].any()     # the wall rule actually fired
    assert (wall["tri_indices"] == owall["tri_indices"]).mean() > 0.999                   # rounding of the position may flip a tie
    assert np.abs(out.astype(np.float64) - oout).max() <= TOL[dt] * h
    # conv = max_i |F_i| s_i: with steps of half a spacing the largest force belongs to a point in a violent rearrangement,
    # where Float32 rounding differences of the force sum (fused vs unfused multiply-adds) show at the 1e-3 level
    np.testing.assert_allclose(conv, oconv, rtol=1e-2 if dt == np.float32 else 1e-9)
    assert oracle.mesh_isinside(sph, out[~is_bnd]).all()                                   # test/repel.jl:31-37


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_repel_deposit_matches_oracle(ctx, oracle, pkg, dt):
    """deposit_ratio > 0 (src/repel.jl:161-168, 328, 483-514): same conversions, landing triangles and positions as the
    serial CPU restatement; cv_target stop before any deposit; deposition needs the mesh-wall method."""
    from test_oracle_golden import _deposit_problem
    sph, snap, is_bnd = _deposit_problem(pkg, np.random.default_rng(9), dt)
    nb = int(is_bnd.sum())
    h = 0.13
    sp, _ = ctx.make_spacing("constant", h)
    osp, _ = oracle.make_spacing("constant", h)
    kw = dict(max_iters=12, tol=0.0, stall_after=0, alpha_lo=h / 2000, alpha_max=h / 2, mesh=sph, is_bnd=is_bnd, deposit_ratio=0.5)
    out, conv, res, _ = ctx.repel(snap, 0, sp, ctx.make_force("clipped", dt(0.2)), **kw)
    wall = ctx.last_wall
    oout, oconv, ores, _ = oracle.repel(snap, 0, osp, oracle.make_force("clipped", dt(0.2)), **kw)
    owall = oracle.repel.last_wall
    assert wall["is_bnd"].sum() > nb                                                          # the boundary grew (test/repel.jl:349)
    assert np.array_equal(wall["is_bnd"], owall["is_bnd"]) and np.array_equal(wall["escaped"], owall["escaped"])
    assert (wall["tri_indices"] == owall["tri_indices"]).mean() > 0.999
    # the wall rule is discontinuous (nearest triangle, inside / outside): a boundary point that sits on an edge of the
    # mesh lands on either of two triangles depending on the last bit of its proposal, which moves it by ~1e-7 h; the
    # north star's 1e-6 s is the tolerance of the smooth (identity-wall) sweep, tested in test_repel_10_iterations_*
    assert np.abs(out.astype(np.float64) - oout).max() <= (1e-5 if dt == np.float64 else 1e-3) * h
    np.testing.assert_allclose(conv, oconv, rtol=1e-2 if dt == np.float32 else 1e-7)
    stopped, _, sres, _ = ctx.repel(snap, 0, sp, ctx.make_force("clipped", dt(0.2)), cv_target=10.0, **kw)
    assert sres["stop_reason"] == "cv_target" and np.array_equal(stopped, snap) and ctx.last_wall["is_bnd"].sum() == nb
    with pytest.raises(pkg.WtpArgumentError):                                                 # not the mesh-wall method
        ctx.repel(snap, nb, sp, ctx.make_force("clipped", dt(0.2)), max_iters=2, alpha_lo=h / 2000, alpha_max=h / 20, deposit_ratio=0.5)


def test_repel_octree_api(ctx, pkg, oracle):
    sph, snap, is_bnd = _wall_problem(pkg, np.random.default_rng(22), np.float64, n_vol=2500, sub=2)
    nb = int(is_bnd.sum())
    normals = sph.face.copy()
    surf = pkg.PointSurface(snap[:nb], normals, np.full(nb, 0.01))
    cloud = pkg.PointCloud(pkg.PointBoundary({"wall": surf}), snap[nb:])
    conv = []
    out = pkg.repel(cloud, pkg.ConstantSpacing(0.11), sph, max_iters=6, stall_after=0, tol=0.0, convergence=conv, ctx=ctx)
    assert len(out) == len(cloud) and len(conv) == 6                  # total point count preserved (test/repel.jl:31-33)
    assert list(out.boundary.surfaces) == ["boundary"] and len(out.boundary) == nb          # _reconstruct_cloud :624-626
    assert pkg.isinside(out.volume.points, sph, ctx=ctx).all()                              # :35-37
    n = out.boundary.surfaces["boundary"].normals
    np.testing.assert_allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-12)
    dep = pkg.repel(cloud, pkg.ConstantSpacing(0.11), sph, max_iters=6, stall_after=0, tol=0.0, deposit_ratio=0.5, alpha=0.05, ctx=ctx)
    assert len(dep) == len(cloud) and len(dep.boundary) >= nb                              # conversions conserve the total (test/repel.jl:348)
    a = dep.boundary.surfaces["boundary"].areas
    assert np.allclose(a[:nb], 0.01) and np.allclose(a[nb:], 0.11 ** 2)                    # deposited points get spacing^2 (:617)
    with pytest.raises(TypeError):
        pkg.repel(cloud, pkg.ConstantSpacing(0.11), deposit_ratio=0.5, ctx=ctx)             # keyword of the octree method only
    with pytest.raises(TypeError):
        pkg.repel(pkg.PointCloud(snap[:50, :2], snap[50:200, :2]), pkg.ConstantSpacing(0.1), sph, ctx=ctx)


# -------------------------------------- isinside(points, cloud) (src/isinside.jl)
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_isinside_greens_matches_oracle(ctx, oracle, pkg, dt):
    rng = np.random.default_rng(31)
    sph = pkg.icosphere_mesh(3)
    tri = sph.triangles.reshape(-1, 3, 3)
    c, n = tri.mean(1).astype(dt), sph.face.astype(dt)
    a = (0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)).astype(dt)
    q = (rng.random((20000, 3)) * 3 - 1.5).astype(dt)
    f, g = ctx.isinside(q, c, n, a, sums=True)
    of, og = oracle.isinside(q, c, n, a)
    tol = 2e-4 if dt == np.float32 else 1e-11                        # same terms, different summation grouping (tiles of 256)
    np.testing.assert_allclose(g, og, rtol=tol, atol=tol)
    clear = np.abs(og + 2 * np.pi) > 1e-2
    assert np.array_equal(f[clear], of[clear]) and clear.mean() > 0.99
    r = np.linalg.norm(q.astype(np.float64), axis=1)
    assert f[r < 0.93].all() and not f[r > 1.05].any()


def test_isinside_polygon_known_answers_and_errors(ctx, pkg, oracle):
    sq = np.array([[0.0, 0], [1, 0], [1, 1], [0, 1]])                 # test/isinside.jl:1-16
    q = np.array([[0.5, 0.5], [0.5, 1.5], [0.5, 1 + np.finfo(float).eps], [1.5, 0.5], [0.5, -0.5]])
    assert pkg.isinside(q, pkg.PointSurface(sq), ctx=ctx).tolist() == [True, False, False, False, False]
    cloud = pkg.PointCloud(pkg.PointBoundary(sq * 3))                  # :29-40
    assert pkg.isinside(np.array([[1.5, 1.5], [4.0, 1.5], [1.5, -1.0]]), cloud, ctx=ctx).tolist() == [True, False, False]
    with pytest.raises(pkg.WtpArgumentError):                          # unordered points (:54-57)
        pkg.isinside(np.array([[0.5, 0.5]]), pkg.PointSurface(np.array([[0.0, 0], [1, 1], [1, 0], [0, 1]])), ctx=ctx)
    with pytest.raises(pkg.WtpArgumentError):                          # too few points (:59-62)
        pkg.isinside(np.array([[0.5, 0.5]]), pkg.PointSurface(np.array([[0.0, 0], [1, 0]])), ctx=ctx)
    rng = np.random.default_rng(32)
    th = np.sort(rng.random(500)) * 2 * np.pi
    poly = np.stack([(1 + 0.3 * np.sin(5 * th)) * np.cos(th), (1 + 0.3 * np.sin(5 * th)) * np.sin(th)], 1)   # non-convex star
    qq = rng.random((20000, 2)) * 3 - 1.5
    f, w = ctx.isinside(qq, poly, sums=True)
    of, ow = oracle.isinside(qq, poly)
    np.testing.assert_allclose(w, ow, atol=1e-9)
    assert np.array_equal(f, of) and 0.2 < f.mean() < 0.6


def test_repel_survivor_filter(ctx, pkg, oracle):
    """repel(cloud, spacing) filters the moved points with isinside(x, cloud) (src/repel.jl:90)."""
    rng = np.random.default_rng(33)
    sph = pkg.icosphere_mesh(2)
    tri = sph.triangles.reshape(-1, 3, 3)
    c = tri.mean(1)
    a = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    vol = rng.normal(size=(3000, 3)); vol *= (1.1 * rng.random((3000, 1)) ** (1 / 3)) / np.linalg.norm(vol, axis=1, keepdims=True)   # some start outside
    cloud = pkg.PointCloud(pkg.PointBoundary({"wall": pkg.PointSurface(c, sph.face, a)}), vol)
    out = pkg.repel(cloud, pkg.ConstantSpacing(0.12), max_iters=3, stall_after=0, tol=0.0, ctx=ctx)   # the default: filter, like the reference
    kept = out.volume.points
    assert len(pkg.repel(cloud, pkg.ConstantSpacing(0.12), max_iters=3, stall_after=0, tol=0.0, isinside=False, ctx=ctx).volume) == len(vol)
    assert 0 < len(kept) < len(vol)
    assert pkg.isinside(kept, cloud, ctx=ctx).all()
    assert (np.linalg.norm(kept, axis=1) < 1.02).all()


# ------------------------------------------ spacing_metrics / spacing_fidelity_metrics
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("D", [2, 3])
def test_spacing_metrics_match_oracle(ctx, oracle, pkg, dt, D):
    rng = np.random.default_rng(40 + D)
    p = rng.random((30000, D)).astype(dt)
    h = len(p) ** (-1.0 / D)
    bnd = np.ascontiguousarray(p[:2000])
    for args in (("constant", h), ("boundary_layer", 0.6 * h, 1.4 * h, 0.3, bnd)):
        sp, k1 = ctx.make_spacing(*args); osp, k2 = oracle.make_spacing(*args)
        a, b = ctx.spacing_metrics(p, sp, 20), oracle.spacing_metrics(p, osp, 20)
        tol = 2e-5 if dt == np.float32 else 1e-11
        for key in b:
            np.testing.assert_allclose(a[key], b[key], rtol=tol, err_msg=key)
        a, b = ctx.spacing_fidelity_metrics(p, sp, 30, 1.4), oracle.spacing_fidelity_metrics(p, osp, 30, 1.4)
        for key in b:
            np.testing.assert_allclose(a[key], b[key], rtol=tol, err_msg=key)
    api = pkg.spacing_fidelity_metrics(pkg.PointCloud(p[:100], p[100:]), pkg.ConstantSpacing(dt(h)), ctx=ctx)
    assert set(api) == {"mean_dnn_h", "cv", "p05", "p50", "p95", "coordination", "k", "coord_radius"} and api["p05"] <= api["p50"] <= api["p95"]
    api = pkg.spacing_metrics(pkg.PointCloud(p[:100], p[100:]), pkg.ConstantSpacing(dt(h)), ctx=ctx)
    assert set(api) == {"max_error", "mean_error", "std_error", "k"} and api["max_error"] >= api["mean_error"] >= 0


def test_repel_kick_after(ctx, pkg):
    """_maybe_kick! (src/repel.jl:415-433; test/repel.jl:399-432): the frozen closest pair is kicked by s/10 in a random
    direction. The random stream is the library's own, so only the semantics are checked: same seed -> same result,
    kick_after = 1 kicks every iteration and moves exactly one point per iteration away from the no-kick trajectory."""
    rng = np.random.default_rng(60)
    for D in (2, 3):
        snap = rng.random((3000, D))
        h = 3000 ** (-1.0 / D)
        sp, _ = ctx.make_spacing("constant", h)
        f = ctx.make_force("clipped", 0.2)
        kw = dict(max_iters=1, tol=0.0, stall_after=0, alpha_lo=h / 2000, alpha_max=h / 20)
        base, _, _, _ = ctx.repel(snap, 300, sp, f, **kw)
        a, conv, res, _ = ctx.repel(snap, 300, sp, f, kick_after=1, kick_seed=7, **kw)
        b, _, _, _ = ctx.repel(snap, 300, sp, f, kick_after=1, kick_seed=7, **kw)
        c, _, _, _ = ctx.repel(snap, 300, sp, f, kick_after=1, kick_seed=8, **kw)
        assert np.array_equal(a, b) and not np.array_equal(a, c)
        moved = np.flatnonzero((a != base).any(axis=1))
        assert len(moved) == 1 and moved[0] >= 300                                  # one movable point, never the fixed wall
        np.testing.assert_allclose(np.linalg.norm(a[moved[0]] - base[moved[0]]), h / 10, rtol=1e-9)
        assert res["iters"] == 1 and np.isfinite(a).all()
        long_run, conv, res, _ = ctx.repel(snap, 300, sp, f, kick_after=5, kick_seed=1, max_iters=30, tol=0.0, stall_after=0,
                                           alpha_lo=h / 2000, alpha_max=h / 20)
        assert res["iters"] == 30 and np.isfinite(long_run).all() and np.array_equal(long_run[:300], snap[:300])
    cloud = pkg.PointCloud(snap[:300], snap[300:])
    out = pkg.repel(cloud, pkg.ConstantSpacing(h), max_iters=3, kick_after=1, isinside=False, ctx=ctx)      # test/repel.jl:417-432
    assert len(out) == len(cloud)


# ------------------------------------------------------------------- cull
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_cull_mask(ctx, oracle, pkg, dt):
    pts = np.array([[0.0, 0, 0], [1, 0, 0], [2, 0, 0], [2.01, 0, 0], [3, 0, 0]], dtype=dt)      # test/repel.jl:301-325
    keep = ctx.cull_mask(pts, np.ones(5, dtype=dt), 0.5)
    assert keep.tolist() == [True, True, True, False, True]
    assert ctx.cull_mask(pts, np.ones(5, dtype=dt), 0.0).all()
    cpts = np.concatenate([pts, np.array([[10.0 + 1.0e-3 * i, 0, 0] for i in range(1, 13)], dtype=dt)])
    ck = ctx.cull_mask(cpts, np.ones(len(cpts), dtype=dt), 0.5)
    assert ck[5:].sum() == 1 and ck[5]
    rng = np.random.default_rng(50)
    for D in (2, 3):
        p = rng.random((40000, D)).astype(dt)
        s = (0.5 + rng.random(len(p))).astype(dt) * len(p) ** (-1.0 / D)                       # variable spacings
        a, b = ctx.cull_mask(p, s, 0.6), oracle.cull_mask(p, s, 0.6)
        assert np.array_equal(a, b) and 0 < (~a).sum() < len(p) // 2
    cloud = pkg.PointCloud(rng.random((200, 3)).astype(dt), rng.random((5000, 3)).astype(dt))
    h = 5000 ** (-1 / 3)
    out = pkg.repel(cloud, pkg.ConstantSpacing(dt(h)), max_iters=2, stall_after=0, tol=0.0, cull_ratio=0.6, isinside=False, ctx=ctx)
    assert len(out.volume) < 5000                                                               # test/repel.jl:375-397: separation guarantee
    m = ctx.metrics(out.volume.points, 2)
    assert m["separation"] >= 0.6 * h * (1 - 1e-6)


# ------------------------------------------------------- BASELINE sizes
def _brute_rows(pts, qi, k):
    """Canonical (d2, index) brute force for a few queries in the input precision (no FMA in numpy)."""
    out = np.empty((len(qi), k), dtype=np.int64)
    for a, i in enumerate(qi):
        d = pts - pts[i]
        d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]
        if pts.shape[1] == 3:
            d2 = d2 + d[:, 2] * d[:, 2]
        cand = np.argpartition(d2, k + 8)[:k + 9]
        order = sorted(cand.tolist(), key=lambda j: (d2[j], j))
        out[a] = np.array(order[1:k + 1]) + 1
    return out


def test_knn_full_size_properties(ctx):
    """BASELINE sizes (1 M and 10 M, k = 21): sampled brute-force agreement, self exclusion,
    ascending distances, and symmetry of the nearest-neighbour relation's distances."""
    rng = np.random.Generator(np.random.Philox(key=0x57545031))
    for n in (1_000_000, 10_000_000):
        pts = rng.random((n, 3)).astype(np.float32)
        idx, dist = ctx.knn(pts, 21, dists=True)
        qi = rng.integers(0, n, 64)
        assert np.array_equal(idx[qi], _brute_rows(pts, qi, 21))
        assert (np.diff(dist, axis=1) >= 0).all()
        assert (idx != np.arange(1, n + 1)[:, None]).all() and idx.min() >= 1 and idx.max() <= n
        d1 = np.sqrt(((pts - pts[idx[:, 0] - 1]) ** 2).sum(1, dtype=np.float64))
        np.testing.assert_allclose(dist[:, 0], d1, rtol=1e-5)
        del idx, dist


def test_radius_full_size_properties(ctx):
    """Config #4 shape (2-D, CSR): 2 M points, r = 2.5 h; sampled brute force + symmetry of the graph."""
    rng = np.random.Generator(np.random.Philox(key=0x57545034))
    n = 2_000_000
    pts = rng.random((n, 2))
    h = n ** -0.5
    off, ind = ctx.radius(pts, 2.5 * h)
    assert off[0] == 0 and (np.diff(off) >= 0).all() and off[-1] == ind.size
    for i in rng.integers(0, n, 32):
        d = pts - pts[i]
        d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]
        want = np.flatnonzero(d2 <= (2.5 * h) * (2.5 * h))
        assert ind[off[i]:off[i + 1]].tolist() == [j + 1 for j in want.tolist() if j != i]
    src = np.repeat(np.arange(1, n + 1), np.diff(off))               # an undirected graph: edge count is symmetric
    assert np.array_equal(np.bincount(src, minlength=n + 1), np.bincount(ind, minlength=n + 1))


def test_repel_large_invariants(ctx):
    """2 M points (config #3 size), 5 iterations: wall untouched, conv finite and decreasing overall,
    closest-pair distance does not collapse, displacement capped at one spacing per sweep."""
    rng = np.random.Generator(np.random.Philox(key=0x57545033))
    n, nf = 2_000_000, 100_000
    snap = rng.random((n, 3)).astype(np.float32)
    h = np.float32(n ** (-1 / 3))
    sp, _ = ctx.make_spacing("constant", h)
    out, conv, res, _ = ctx.repel(snap, nf, sp, ctx.make_force("clipped", 0.2), max_iters=5, tol=0.0, stall_after=0,
                                  alpha_lo=h / 2000, alpha_max=h / 20)
    assert res["iters"] == 5 and np.isfinite(conv).all() and (conv >= 0).all() and conv[-1] < conv[0]
    assert np.array_equal(out[:nf], snap[:nf])
    assert np.sqrt(((out - snap) ** 2).sum(1)).max() <= 5 * h * 1.0001
