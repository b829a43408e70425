"""ctypes front end of the CPU oracle (oracle/wtp_oracle.cpp).

TEST INFRASTRUCTURE ONLY — "parity unpinned" for neighbour identities and repel
trajectories (see the header of wtp_oracle.cpp). May be imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by
the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libwtp_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "wtp_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "wtp_cuda.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


class Force(C.Structure):
    _fields_ = [("kind", C.c_int32), ("beta", C.c_double), ("u0", C.c_double), ("gamma", C.c_double)]


class Spacing(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_double), ("b", C.c_double), ("c", C.c_double),
                ("bnd_pts", C.c_void_p), ("n_bnd", C.c_int64)]


class RepelParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("max_iters", C.c_int32), ("rebuild_every", C.c_int32),
                ("stall_after", C.c_int32), ("kick_after", C.c_int32), ("wall", C.c_int32),
                ("want_trace", C.c_int32), ("reserved", C.c_int32),
                ("alpha_lo", C.c_double), ("alpha_max", C.c_double),
                ("tol", C.c_double), ("cv_target", C.c_double),
                ("n_protected", C.c_int64), ("kick_seed", C.c_uint64), ("deposit_ratio", C.c_double)]


class RepelResult(C.Structure):
    _fields_ = [("iters", C.c_int32), ("stop_reason", C.c_int32), ("last_cv", C.c_double)]


class TraceEntry(C.Structure):
    _fields_ = [("r", C.c_double), ("s", C.c_double), ("r_over_s", C.c_double),
                ("idx_a", C.c_int64), ("idx_b", C.c_int64)]


class WallMesh(C.Structure):
    _fields_ = [("triangles", C.c_void_p), ("feature_normals", C.c_void_p), ("n_tri", C.c_int64),
                ("bbox_min", C.c_double * 3), ("bbox_max", C.c_double * 3), ("offset_dist", C.c_double),
                ("is_bnd", C.c_void_p), ("tri_indices", C.c_void_p), ("escaped", C.c_void_p)]


def make_wall(triangles, feature_normals, bbox_min, bbox_max, offset_dist, is_bnd=None, tri_indices=None, escaped=None):
    """WallMesh from the flattened TriangleIndex arrays (n x 9, n x 21). Returns (struct, keepalive)."""
    triangles, feature_normals = np.ascontiguousarray(triangles), np.ascontiguousarray(feature_normals)
    w = WallMesh()
    w.triangles, w.feature_normals, w.n_tri = triangles.ctypes.data, feature_normals.ctypes.data, triangles.shape[0]
    for d in range(3):
        w.bbox_min[d], w.bbox_max[d] = float(bbox_min[d]), float(bbox_max[d])
    w.offset_dist = float(offset_dist)
    w.is_bnd = is_bnd.ctypes.data if is_bnd is not None else None
    w.tri_indices = tri_indices.ctypes.data if tri_indices is not None else None
    w.escaped = escaped.ctypes.data if escaped is not None else None
    return w, (triangles, feature_normals, is_bnd, tri_indices, escaped)


def wall_from(mesh, dtype, is_bnd=None, tri_indices=None, escaped=None):
    """Same, from any object with the TriangleIndex arrays (e.g. the host mirror's TriangleOctree)."""
    m = mesh.astype(dtype)
    return make_wall(m.triangles, m.feature_normals, m.bbox_min, m.bbox_max, m.offset_dist, is_bnd, tri_indices, escaped)


def mesh_isinside(mesh, pts, *, threads=0):
    pts = _pts(pts)
    w, keep = wall_from(mesh, pts.dtype)
    out = np.zeros(pts.shape[0], dtype=np.uint8)
    getattr(lib(), "wtpo_mesh_isinside_" + _sfx(pts.dtype))(C.byref(w), pts.ctypes.data_as(C.c_void_p), C.c_int64(pts.shape[0]),
                                                            C.c_int32(threads), out.ctypes.data_as(C.c_void_p))
    return out.astype(bool)


def mesh_project(mesh, pts, *, threads=0):
    pts = _pts(pts)
    w, keep = wall_from(mesh, pts.dtype)
    out, tri = np.empty_like(pts), np.zeros(pts.shape[0], dtype=np.int64)
    getattr(lib(), "wtpo_mesh_project_" + _sfx(pts.dtype))(C.byref(w), pts.ctypes.data_as(C.c_void_p), C.c_int64(pts.shape[0]),
                                                           C.c_int32(threads), out.ctypes.data_as(C.c_void_p), tri.ctypes.data_as(C.c_void_p))
    return out, tri


def isinside(pts, bnd_pts, bnd_normals=None, bnd_areas=None, *, threads=0):
    """isinside(points, cloud) (src/isinside.jl): 3-D Green's function / 2-D winding number. Returns (flags, sums)."""
    pts = _pts(pts)
    d = pts.shape[1]
    bx = np.ascontiguousarray(bnd_pts, dtype=pts.dtype)
    bn = np.ascontiguousarray(bnd_normals, dtype=pts.dtype) if bnd_normals is not None else None
    ba = np.ascontiguousarray(bnd_areas, dtype=pts.dtype) if bnd_areas is not None else None
    out = np.zeros(pts.shape[0], dtype=np.uint8)
    sums = np.zeros(pts.shape[0], dtype=pts.dtype)
    vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    getattr(lib(), "wtpo_isinside_" + _sfx(pts.dtype))(vp(pts), C.c_int64(pts.shape[0]), C.c_int32(d), vp(bx), vp(bn), vp(ba),
                                                        C.c_int64(bx.shape[0]), C.c_int32(threads), vp(out), vp(sums))
    return out.astype(bool), sums


class CloudMetrics(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("avg", "std", "max", "min", "separation", "fill", "mesh_ratio")]


FORCE_KINDS = {"inverse": 0, "equilibrium": 1, "clipped": 2, "strong": 3}
SPACING_KINDS = {"constant": 0, "loglike": 1, "boundary_layer": 2}
STOP_REASONS = {0: "max_iters", 1: "tol", 2: "cv_target", 3: "stall"}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def _pts(pts):
    pts = np.ascontiguousarray(pts)
    assert pts.ndim == 2 and pts.shape[1] in (2, 3), pts.shape
    return pts


def max_threads() -> int:
    """Threads the oracle uses by default (OpenMP's view: OMP_NUM_THREADS, which torchrun sets to 1)."""
    return int(lib().wtpo_max_threads())


def host_threads() -> int:
    """The cores this process may run on (sched_getaffinity): what bench.py passes as `threads=` so that the CPU arm
    uses the same number of threads whatever launcher exported OMP_NUM_THREADS."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


def knn(pts, k, *, drop_first=True, algo="kdtree", threads=0, dists=False):
    """_build_knn_neighbors (drop_first) / search+searchdists (not drop_first)."""
    pts = _pts(pts)
    n, d = pts.shape
    idx = np.empty((n, k), dtype=np.int64)
    dist = np.empty((n, k), dtype=pts.dtype) if dists else None
    fn = getattr(lib(), "wtpo_knn_" + _sfx(pts.dtype))
    rc = fn(pts.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_int32(d), C.c_int32(k),
            C.c_int32(1 if drop_first else 0), C.c_int32(0 if algo == "brute" else 1),
            C.c_int32(threads), idx.ctypes.data_as(C.c_void_p),
            dist.ctypes.data_as(C.c_void_p) if dists else None)
    if rc != 0:
        raise ValueError(f"oracle knn failed with status {rc}")
    return (idx, dist) if dists else idx


def radius(pts, r, *, threads=0):
    """_build_radius_neighbors as CSR (offsets 0-based, indices 1-based ascending)."""
    pts = _pts(pts)
    n, d = pts.shape
    offsets = np.zeros(n + 1, dtype=np.int64)
    sfx = _sfx(pts.dtype)
    fn = getattr(lib(), "wtpo_radius_" + sfx)
    rr = C.c_float(r) if sfx == "f32" else C.c_double(r)
    rc = fn(pts.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_int32(d), rr, C.c_int32(threads),
            offsets.ctypes.data_as(C.c_void_p), None)
    assert rc == 0, rc
    indices = np.empty(int(offsets[-1]), dtype=np.int64)
    rc = fn(pts.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_int32(d), rr, C.c_int32(threads),
            offsets.ctypes.data_as(C.c_void_p), indices.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    return offsets, indices


def make_force(kind="clipped", beta=0.2, u0=1.0, gamma=3.0) -> Force:
    return Force(FORCE_KINDS[kind], float(beta), float(u0), float(gamma))


def make_spacing(kind="constant", a=0.0, b=0.0, c=0.0, bnd_pts=None):
    """Returns (Spacing, keepalive)."""
    if bnd_pts is not None:
        bnd_pts = np.ascontiguousarray(bnd_pts)
        return Spacing(SPACING_KINDS[kind], float(a), float(b), float(c),
                       bnd_pts.ctypes.data, bnd_pts.shape[0]), bnd_pts
    return Spacing(SPACING_KINDS[kind], float(a), float(b), float(c), None, 0), None


def force(f: Force, u):
    u = np.ascontiguousarray(u)
    out = np.empty_like(u)
    getattr(lib(), "wtpo_force_" + _sfx(u.dtype))(C.byref(f), u.ctypes.data_as(C.c_void_p),
                                                  C.c_int64(u.size), out.ctypes.data_as(C.c_void_p))
    return out


def spacing_eval(sp: Spacing, pts):
    pts = _pts(pts)
    out = np.empty(pts.shape[0], dtype=pts.dtype)
    rc = getattr(lib(), "wtpo_spacing_" + _sfx(pts.dtype))(
        C.byref(sp), pts.ctypes.data_as(C.c_void_p), C.c_int64(pts.shape[0]), C.c_int32(pts.shape[1]),
        out.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    return out


def repel(snap, n_fixed, sp: Spacing, f: Force, *, k=21, max_iters=1000, tol=1e-6, rebuild_every=1,
          stall_after=50, cv_target=0.0, alpha_lo, alpha_max, kick_after=0, trace=False, threads=0, mesh=None, is_bnd=None,
          deposit_ratio=0.0, kick_seed=0, cv_in_double=False):
    """_relax! on snap = [fixed head; movable tail]. Returns (new_snap, conv, result dict, trace)."""
    snap = np.array(_pts(snap), copy=True)
    n_all, d = snap.shape
    n_move = n_all - n_fixed
    # cv_in_double: NOT the reference — accumulate the d_NN/s sums of the stop test in double like the device does (its
    # documented deviation), to isolate what the accumulation precision does to a Float32 stall_after run
    prm = RepelParams(k, max_iters, rebuild_every, stall_after, kick_after, 1 if mesh is not None else 0, 1 if trace else 0,
                      1 if cv_in_double else 0,
                      float(alpha_lo), float(alpha_max), float(tol), float(cv_target), 0, int(kick_seed), float(deposit_ratio))
    conv = np.zeros(max(max_iters, 1), dtype=snap.dtype)
    tr = (TraceEntry * max(max_iters, 1))() if trace else None
    res = RepelResult()
    wall, keep = None, None
    repel.last_wall = None
    if mesh is not None:
        flags = np.ascontiguousarray(is_bnd, dtype=np.uint8)
        tri_idx, esc = np.zeros(n_move, dtype=np.int64), np.zeros(n_move, dtype=np.uint8)
        w, keep = wall_from(mesh, snap.dtype, flags, tri_idx, esc)
        wall = C.byref(w)
        repel.last_wall = dict(tri_indices=tri_idx, escaped=esc, is_bnd=flags)
    rc = getattr(lib(), "wtpo_repel_" + _sfx(snap.dtype))(
        snap.ctypes.data_as(C.c_void_p), C.c_int64(n_fixed), C.c_int64(n_move), C.c_int32(d),
        C.byref(sp), C.byref(f), C.byref(prm), wall, conv.ctypes.data_as(C.c_void_p),
        tr, C.byref(res), C.c_int32(threads))
    if rc != 0:
        raise ValueError(f"oracle repel failed with status {rc}")
    out_tr = None
    if trace:
        out_tr = [dict(iteration=i + 1, r=tr[i].r, s=tr[i].s, r_over_s=tr[i].r_over_s,
                       idx_a=tr[i].idx_a, idx_b=tr[i].idx_b) for i in range(res.iters)]
    return snap, conv[:res.iters].copy(), dict(iters=res.iters, stop_reason=STOP_REASONS[res.stop_reason],
                                               last_cv=res.last_cv), out_tr


def metrics(pts, k=20, *, threads=0):
    pts = _pts(pts)
    out = CloudMetrics()
    rc = getattr(lib(), "wtpo_metrics_" + _sfx(pts.dtype))(
        pts.ctypes.data_as(C.c_void_p), C.c_int64(pts.shape[0]), C.c_int32(pts.shape[1]), C.c_int32(k),
        C.c_int32(threads), C.byref(out))
    if rc != 0:
        raise ValueError(f"oracle metrics failed with status {rc}")
    return {n: getattr(out, n) for n, _ in CloudMetrics._fields_}


def cull_mask(pts, spacings, ratio):
    """_near_duplicate_keep_mask(pts, spacings, ratio) (src/repel.jl:565-580) -> bool keep mask."""
    pts = _pts(pts)
    sp = np.ascontiguousarray(spacings, dtype=pts.dtype)
    keep = np.ones(pts.shape[0], dtype=np.uint8)
    rc = getattr(lib(), "wtpo_cull_mask_" + _sfx(pts.dtype))(pts.ctypes.data_as(C.c_void_p), C.c_int64(pts.shape[0]), C.c_int32(pts.shape[1]),
                                                             sp.ctypes.data_as(C.c_void_p), C.c_double(ratio), keep.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    return keep.astype(bool)


class SpacingMetrics(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("max_error", "mean_error", "std_error")]


class SpacingFidelity(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("mean_dnn_h", "cv", "p05", "p50", "p95", "coordination")]


def spacing_metrics(pts, sp: Spacing, k=20, *, threads=0):
    pts = _pts(pts)
    out = SpacingMetrics()
    rc = getattr(lib(), "wtpo_spacing_metrics_" + _sfx(pts.dtype))(pts.ctypes.data_as(C.c_void_p), C.c_int64(pts.shape[0]), C.c_int32(pts.shape[1]),
                                                                   C.c_int32(k), C.byref(sp), C.c_int32(threads), C.byref(out))
    if rc != 0:
        raise ValueError(f"oracle spacing_metrics failed with status {rc}")
    return {n: getattr(out, n) for n, _ in SpacingMetrics._fields_}


def spacing_fidelity_metrics(pts, sp: Spacing, k=30, coord_radius=1.4, *, threads=0):
    pts = _pts(pts)
    out = SpacingFidelity()
    rc = getattr(lib(), "wtpo_spacing_fidelity_" + _sfx(pts.dtype))(pts.ctypes.data_as(C.c_void_p), C.c_int64(pts.shape[0]), C.c_int32(pts.shape[1]),
                                                                    C.c_int32(k), C.c_double(coord_radius), C.byref(sp), C.c_int32(threads), C.byref(out))
    if rc != 0:
        raise ValueError(f"oracle spacing_fidelity_metrics failed with status {rc}")
    return {n: getattr(out, n) for n, _ in SpacingFidelity._fields_}


def normals(pts, k=5, *, threads=0):
    """compute_normals(points; k) (src/normals.jl:9-44): unit PCA normals, N x D, first nonzero component positive."""
    pts = _pts(pts)
    out = np.empty_like(pts)
    rc = getattr(lib(), "wtpo_normals_" + _sfx(pts.dtype))(pts.ctypes.data_as(C.c_void_p), C.c_int64(pts.shape[0]), C.c_int32(pts.shape[1]),
                                                          C.c_int32(k), C.c_int32(threads), out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise ValueError(f"oracle normals failed with status {rc}")
    return out


def gradient_limit(centers, h0, g, k=12, tol=1.0e-3, max_sweeps=2000, *, threads=0):
    """_gradient_limit_field (src/discretization/algorithms/octree.jl:677-717) on the leaf centres -> (h, sweeps)."""
    centers = _pts(centers)
    h0 = np.ascontiguousarray(h0, dtype=centers.dtype)
    out = np.empty_like(h0)
    sweeps = C.c_int32(0)
    sfx = _sfx(centers.dtype)
    gg = C.c_float(g) if sfx == "f32" else C.c_double(g)
    rc = getattr(lib(), "wtpo_gradient_limit_" + sfx)(centers.ctypes.data_as(C.c_void_p), C.c_int64(centers.shape[0]), C.c_int32(centers.shape[1]),
                                                     h0.ctypes.data_as(C.c_void_p), gg, C.c_int32(k), C.c_double(tol), C.c_int32(max_sweeps),
                                                     C.c_int32(threads), out.ctypes.data_as(C.c_void_p), C.byref(sweeps))
    if rc != 0:
        raise ValueError(f"oracle gradient_limit failed with status {rc}")
    return out, int(sweeps.value)


def closest_point_on_triangle(p, a, b, c):
    p, a, b, c = (np.ascontiguousarray(x) for x in (p, a, b, c))
    out = np.empty(3, dtype=p.dtype)
    feature = getattr(lib(), "wtpo_closest_point_on_triangle_" + _sfx(p.dtype))(
        *(x.ctypes.data_as(C.c_void_p) for x in (p, a, b, c)), out.ctypes.data_as(C.c_void_p))
    closest_point_on_triangle.last_feature = int(feature)
    return out
