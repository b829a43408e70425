"""Other consumers of the k-NN index: compute_normals (src/normals.jl:9-44, 65-70) and
_gradient_limit_field (src/discretization/algorithms/octree.jl:677-717).

CPU: the oracle's restatements against independent implementations (numpy.linalg.eigh over brute-force neighbourhoods;
a plain-Python Bellman sweep) and closed forms (a plane, a sphere, a single source whose envelope is h0 + g d).
GPU (-m gpu): the device entry points against the oracle — the gradient-limit field bit for bit (same neighbour lists,
same multiply-then-add, same stop sweep), the normals to rounding up to sign."""
import numpy as np
import pytest


def _np_normals(pts, k):
    out = np.empty_like(pts)
    for i in range(len(pts)):
        d2 = ((pts - pts[i]) ** 2).sum(1)
        nb = np.lexsort((np.arange(len(pts)), d2))[:k]
        w, q = np.linalg.eigh(np.cov(pts[nb].astype(np.float64).T))
        v = q[:, 0]
        out[i] = v * (1 if v[np.flatnonzero(v)[0]] > 0 else -1)
    return out


@pytest.mark.parametrize("D", [2, 3])
def test_oracle_normals_match_numpy_and_closed_forms(oracle, D):
    rng = np.random.default_rng(5 + D)
    pts = rng.random((400, D))
    a, b = oracle.normals(pts, 7), _np_normals(pts, 7)
    assert np.abs((a * b).sum(1)).min() > 1 - 1e-9                    # same direction up to sign
    np.testing.assert_allclose(np.linalg.norm(a, axis=1), 1.0, atol=1e-12)
    if D == 3:
        plane = rng.random((500, 3)); plane[:, 2] = 0.25 + 1e-3 * plane[:, 0]      # a tilted plane: every normal is its normal
        n = oracle.normals(plane, 6)
        want = np.array([-1e-3, 0, 1]) / np.linalg.norm([-1e-3, 0, 1])
        assert np.abs(n @ want).min() > 1 - 1e-9
        sph = rng.normal(size=(4000, 3)); sph /= np.linalg.norm(sph, axis=1, keepdims=True)
        n = oracle.normals(sph, 8)
        assert np.abs((n * sph).sum(1)).min() > 0.97 and np.abs((n * sph).sum(1)).mean() > 0.999   # radial, to the curvature of an 8-point patch
    assert oracle.normals(pts[:4], 50).shape == (4, D)                # k clamped to N (src/normals.jl:16)


def _py_gradient_limit(c, h0, g, k, tol, max_sweeps):
    n = len(c)
    kk = min(k, n)
    nbr, dd = [], []
    for a in range(n):
        d2 = ((c - c[a]) ** 2).sum(1)
        o = np.lexsort((np.arange(n), d2))[:kk]
        nbr.append(o); dd.append(np.sqrt(d2[o]))
    h = h0.copy()
    sweeps = 0
    for _ in range(max_sweeps):
        hn = np.array([min(h[a], (h[nbr[a]] + g * dd[a]).min()) for a in range(n)])
        maxrel = (np.abs(hn - h) / h).max()
        h = hn
        sweeps += 1
        if maxrel < tol:
            break
    return h, sweeps


def test_oracle_gradient_limit_matches_python_and_envelope(oracle):
    rng = np.random.default_rng(8)
    c = rng.random((600, 3))
    h0 = 0.05 + 0.5 * rng.random(600)
    a, sa = oracle.gradient_limit(c, h0, 0.3, k=12, tol=1e-3)
    b, sb = _py_gradient_limit(c, h0, 0.3, 12, 1e-3, 2000)
    assert sa == sb and np.array_equal(a, b)
    assert (a <= h0).all() and (a > 0).all()
    # one fine source in a coarse field: the envelope is min(h0, h_src + g * graph distance) >= h_src + g * |x - x_src|
    g2 = np.stack(np.meshgrid(*[np.arange(20) / 19.0] * 2, indexing="ij"), -1).reshape(-1, 2)
    h0 = np.full(len(g2), 1.0); h0[0] = 0.01
    e, _ = oracle.gradient_limit(g2, h0, 0.5, k=9, tol=1e-6)
    lower = np.minimum(1.0, 0.01 + 0.5 * np.linalg.norm(g2 - g2[0], axis=1))
    assert (e >= lower - 1e-12).all() and e[0] == 0.01
    along = np.flatnonzero(g2[:, 1] == 0)                              # along a lattice line the graph distance is Euclidean
    np.testing.assert_allclose(e[along], lower[along], rtol=1e-12)
    x, s = oracle.gradient_limit(c[:5], np.ones(5), 0.3, k=12)         # k clamped to n; nothing to limit: one sweep
    assert s == 1 and np.array_equal(x, np.ones(5))


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("D", [2, 3])
def test_normals_match_oracle(ctx, oracle, dt, D):
    rng = np.random.default_rng(20 + D)
    if D == 3:                                                        # a wavy sheet: a real surface with noise
        uv = rng.random((60000, 2))
        pts = np.stack([uv[:, 0], uv[:, 1], 0.1 * np.sin(6 * uv[:, 0]) * np.cos(4 * uv[:, 1]) + 1e-4 * rng.normal(size=len(uv))], 1)
    else:
        t = np.sort(rng.random(20000)) * 2 * np.pi
        pts = np.stack([(1 + 0.2 * np.sin(5 * t)) * np.cos(t), (1 + 0.2 * np.sin(5 * t)) * np.sin(t)], 1) + 1e-5 * rng.normal(size=(len(t), 2))
    pts = np.ascontiguousarray(pts.astype(dt))
    for k in (5, 10):
        a, b = ctx.normals(pts, k), oracle.normals(pts, k)
        assert a.shape == b.shape and a.dtype == dt
        dots = np.abs((a.astype(np.float64) * b.astype(np.float64)).sum(1))
        assert dots.min() > 1 - (1e-5 if dt == np.float32 else 1e-12)  # same covariance (same neighbours, same T arithmetic), same eigenvector
        assert np.array_equal(np.sign(a[np.arange(len(a)), np.argmax(a != 0, axis=1)]), np.ones(len(a)))
    assert ctx.normals(pts[:3], 5).shape == (3, D)                    # k > N is clamped


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("D", [2, 3])
def test_gradient_limit_bit_exact(ctx, oracle, dt, D):
    rng = np.random.default_rng(30 + D)
    n = 50000
    c = rng.random((n, D)).astype(dt)
    h0 = (0.002 + 0.2 * rng.random(n) ** 4).astype(dt)                # mostly fine with coarse outliers, plus a few very fine sources
    h0[rng.integers(0, n, 20)] = dt(1e-4)
    for g, k, tol in ((0.3, 12, 1e-3), (0.1, 6, 1e-6)):
        a, sa = ctx.gradient_limit(c, h0, g, k=k, tol=tol)
        b, sb = oracle.gradient_limit(c, h0, g, k=k, tol=tol)
        assert sa == sb and np.array_equal(a, b)
        assert (a <= h0).all() and (a < h0).mean() > 0.2              # the limiter actually did something
    a, sa = ctx.gradient_limit(c, h0, 0.3, k=12, tol=1e-9, max_sweeps=3)
    b, sb = oracle.gradient_limit(c, h0, 0.3, k=12, tol=1e-9, max_sweeps=3)
    assert sa == sb == 3 and np.array_equal(a, b)                     # the sweep cap
