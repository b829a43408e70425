"""Synthetic inputs of SURVEY.md §8(d): U3 (uniform cube), G3 (boundary-layer graded cube),
Q2 (quadtree-graded square). Counter-based Philox streams, key 0x57545031 + config id."""
import numpy as np

KEY = 0x57545031


def _rng(stream):
    return np.random.Generator(np.random.Philox(key=KEY + stream))


def uniform_cube(n, dtype=np.float32, stream=0):
    return _rng(stream).random((n, 3)).astype(dtype)


def _h_of_d(d, h_wall, ratio, delta):
    return h_wall + (ratio - 1.0) * h_wall / (1.0 + np.exp(-(d - delta / 2) / (delta / 6)))


def graded_cube(n_total, dtype=np.float64, ratio=4.0, delta=0.2, stream=3):
    """G3: wall lattice on the 6 faces at h_wall (fixed points = the spacing's boundary set) +
    interior points drawn with density 1/h(d)^3, h = BoundaryLayerSpacing(at_wall=h_wall,
    bulk=ratio*h_wall, layer_thickness=delta). h_wall is solved so that the total is ~n_total.
    Returns (points [wall first], n_wall, h_wall)."""
    dd = np.linspace(0, 0.5, 20001)
    shell = 6.0 * (1 - 2 * dd) ** 2                      # dV/dd of the unit cube at wall distance d

    def count(hw):
        m = int(round(1.0 / hw))
        wall = 6 * (m + 1) ** 2 - 12 * (m + 1) + 8
        return wall + np.trapezoid(shell / _h_of_d(dd, hw, ratio, delta) ** 3, dd)

    lo, hi = 1e-4, 0.5
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        if count(mid) > n_total:
            lo = mid
        else:
            hi = mid
    hw = 0.5 * (lo + hi)
    m = int(round(1.0 / hw))
    g = np.linspace(0.0, 1.0, m + 1)
    a, b = np.meshgrid(g, g, indexing="ij")
    a, b = a.ravel(), b.ravel()
    faces = [np.stack([np.zeros_like(a), a, b], 1), np.stack([np.ones_like(a), a, b], 1),
             np.stack([a, np.zeros_like(a), b], 1), np.stack([a, np.ones_like(a), b], 1),
             np.stack([a, b, np.zeros_like(a)], 1), np.stack([a, b, np.ones_like(a)], 1)]
    wall = np.unique(np.concatenate(faces), axis=0)
    n_int = max(n_total - len(wall), 0)
    rng = _rng(stream)
    out = []
    have = 0
    while have < n_int:
        c = rng.random((max(4 * (n_int - have), 1024), 3)) * (1 - hw) + hw / 2   # keep off the wall itself
        d = np.minimum(c, 1 - c).min(1)
        keep = rng.random(len(c)) < (hw / _h_of_d(d, hw, ratio, delta)) ** 3
        out.append(c[keep])
        have += int(keep.sum())
    interior = np.concatenate(out)[:n_int]
    pts = np.concatenate([wall, interior]).astype(dtype)
    return pts, len(wall), hw


def graded_square(n_total, dtype=np.float64, stream=4):
    """Q2: unit square, 3-level quadtree grading toward the origin corner (h, h/2, h/4),
    jittered lattice per region. Returns (points, h_mid)."""
    # areas: level2 (finest) [0,1/4]^2, level1 [0,1/2]^2 minus that, level0 the rest
    areas = np.array([1 - 0.25, 0.25 - 0.0625, 0.0625])
    dens = np.array([1.0, 4.0, 16.0])
    h = np.sqrt((areas * dens).sum() / n_total)
    rng = _rng(stream)
    pts = []
    for lvl, (lo_, hi_) in enumerate([(0.5, 1.0), (0.25, 0.5), (0.0, 0.25)]):
        hh = h / (2 ** lvl)
        m = int(np.ceil(hi_ / hh))
        g = (np.arange(m) + 0.5) * hh
        x, y = np.meshgrid(g, g, indexing="ij")
        p = np.stack([x.ravel(), y.ravel()], 1)
        p += (rng.random(p.shape) - 0.5) * 0.5 * hh
        mx = np.maximum(p[:, 0], p[:, 1])
        pts.append(p[(mx < hi_) & (mx >= lo_) & (p.min(1) >= 0)])
    return np.concatenate(pts).astype(dtype), h / 2
