"""Host-side mesh index for the wall rule of `repel(cloud, spacing, octree)`.

Mirrors what the reference builds on the host at `TriangleOctree` construction (which is
outside the hot path and stays host code): the `TriangleIndex`
arrays of src/octree/triangle_octree.jl:235-290 — unit face normals, the angle-weighted
vertex pseudonormals and the edge pseudonormals (Bærentzen & Aanæs 2005) keyed by exact
coordinates, and the mesh bounding box. The spatial search structure itself (the octree in
the reference) is replaced by a device BVH + cell-class grid built inside libwtp_cuda.so.
"""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from ._lib import WallMesh


def read_binary_stl(path: str) -> np.ndarray:
    """Triangle soup (n x 3 x 3 float32) of a binary STL file."""
    with open(path, "rb") as f:
        f.read(80)
        n = struct.unpack("<I", f.read(4))[0]
        rec = np.frombuffer(f.read(50 * n), dtype=np.uint8).reshape(n, 50)
    return rec[:, 12:48].copy().view(np.float32).reshape(n, 3, 3)


class TriangleOctree:
    """`TriangleOctree(mesh)` stand-in: TriangleIndex arrays + flattened per-triangle features."""

    def __init__(self, triangles, vertices=None, dtype=np.float64):
        T = np.dtype(dtype).type
        tri = np.asarray(triangles)
        if vertices is not None:                         # indexed mesh (0-based connectivity)
            tri = np.asarray(vertices, dtype=dtype)[tri.astype(np.int64)]
        tri = np.ascontiguousarray(tri, dtype=dtype).reshape(-1, 3, 3)
        n = tri.shape[0]
        if n == 0:
            raise ValueError("Mesh must contain at least one triangle")
        eps = np.finfo(dtype).eps
        a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
        nrm = np.cross(b - a, c - a)
        mag = np.sqrt((nrm * nrm).sum(1))
        face = np.where((mag < eps * 100)[:, None], 0, nrm / np.where(mag == 0, 1, mag)[:, None]).astype(dtype)   # :258-260
        edge, vertex = {}, {}
        zero = np.zeros(3, dtype=dtype)

        def key(p, q):                                   # _edge_key :305-307 (lexicographic)
            p, q = tuple(p.tolist()), tuple(q.tolist())
            return (p, q) if p < q else (q, p)

        def corner(vc, va, vb):                          # _corner_angle :309-315
            u, w = va - vc, vb - vc
            den = np.sqrt(T(u @ u) * T(w @ w))
            if den < eps:
                return T(0)
            return T(np.arccos(np.clip(T(u @ w) / den, T(-1), T(1))))

        for i in range(n):                               # :270-281
            v1, v2, v3 = tri[i]
            nh = face[i]
            for va, vb in ((v1, v2), (v2, v3), (v3, v1)):
                k = key(va, vb)
                edge[k] = edge.get(k, zero) + nh
            for vc, va, vb in ((v1, v2, v3), (v2, v3, v1), (v3, v1, v2)):
                kv = tuple(vc.tolist())
                vertex[kv] = vertex.get(kv, zero) + corner(vc, va, vb) * nh
        feat = np.empty((n, 7, 3), dtype=dtype)          # _feature_pseudonormal :322-337
        for i in range(n):
            v1, v2, v3 = tri[i]
            feat[i, 0] = face[i]
            feat[i, 1] = vertex[tuple(v1.tolist())]
            feat[i, 2] = vertex[tuple(v2.tolist())]
            feat[i, 3] = vertex[tuple(v3.tolist())]
            feat[i, 4] = edge[key(v1, v2)]
            feat[i, 5] = edge[key(v1, v3)]
            feat[i, 6] = edge[key(v2, v3)]
        verts = tri.reshape(-1, 3)
        lo, hi = verts.min(0).astype(dtype), verts.max(0).astype(dtype)    # _compute_bbox_raw :293-304
        eps_val = max(eps * 100, 1.0e-10)
        for d in range(3):
            if lo[d] == hi[d]:
                lo[d] -= eps_val
                hi[d] += eps_val
        self.dtype = np.dtype(dtype)
        self.triangles = np.ascontiguousarray(tri.reshape(n, 9))
        self.face = face
        self.feature_normals = np.ascontiguousarray(feat.reshape(n, 21))
        self.bbox_min, self.bbox_max = lo, hi
        self.offset_dist = float(T(1.0e-6) * np.sqrt(((hi - lo) ** 2).sum()))          # src/repel.jl:150

    def __len__(self):
        return self.triangles.shape[0]

    def signed_volume(self) -> float:                    # _signed_volume :386-393
        t = self.triangles.reshape(-1, 3, 3).astype(np.float64)
        return float((t[:, 0] * np.cross(t[:, 1], t[:, 2])).sum() / 6)

    def astype(self, dtype) -> "TriangleOctree":
        if np.dtype(dtype) == self.dtype:
            return self
        return TriangleOctree(self.triangles.reshape(-1, 3, 3), dtype=dtype)

    def wall(self, is_bnd=None, tri_indices=None, escaped=None):
        """(WallMesh struct, keepalive tuple) for the C ABI."""
        w = WallMesh()
        w.triangles = self.triangles.ctypes.data
        w.feature_normals = self.feature_normals.ctypes.data
        w.n_tri = len(self)
        for d in range(3):
            w.bbox_min[d] = float(self.bbox_min[d])
            w.bbox_max[d] = float(self.bbox_max[d])
        w.offset_dist = self.offset_dist
        w.is_bnd = is_bnd.ctypes.data if is_bnd is not None else None
        w.tri_indices = tri_indices.ctypes.data if tri_indices is not None else None
        w.escaped = escaped.ctypes.data if escaped is not None else None
        return w, (self, is_bnd, tri_indices, escaped)


def unit_cube_mesh(dtype=np.float64) -> TriangleOctree:
    """The reference's test mesh (test/testsetup.jl:31-50): 8 vertices, 12 outward-oriented triangles."""
    v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=dtype)
    t = np.array([[1, 3, 2], [1, 4, 3], [5, 6, 7], [5, 7, 8], [1, 2, 6], [1, 6, 5], [3, 4, 8], [3, 8, 7], [1, 5, 8], [1, 8, 4],
                  [2, 3, 7], [2, 7, 6]]) - 1
    return TriangleOctree(t, v, dtype=dtype)


def cuboid_mesh(lx, ly, lz, dtype=np.float64) -> TriangleOctree:
    """Axis-aligned box [0,lx]x[0,ly]x[0,lz] with the unit cube's connectivity (test/octree_isinside.jl:66-92)."""
    m = unit_cube_mesh(dtype)
    tri = m.triangles.reshape(-1, 3, 3) * np.array([lx, ly, lz], dtype=dtype)
    return TriangleOctree(tri, dtype=dtype)


def icosphere_mesh(subdivisions=3, radius=1.0, center=(0.0, 0.0, 0.0), dtype=np.float64) -> TriangleOctree:
    """Closed, outward-oriented triangulated sphere (20 * 4^subdivisions triangles): a synthetic
    watertight test geometry that needs no file from the reference checkout."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    verts = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdivisions):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    vv = (np.array(verts) * radius + np.array(center, dtype=np.float64)).astype(dtype)
    return TriangleOctree(np.array(f, dtype=np.int64), vv, dtype=dtype)


def torus_mesh(R=1.0, r=0.35, nu=48, nv=24, dtype=np.float64) -> TriangleOctree:
    """Closed, outward-oriented torus (2 * nu * nv triangles): a non-convex watertight test geometry."""
    u = np.arange(nu) * 2 * np.pi / nu
    w = np.arange(nv) * 2 * np.pi / nv
    uu, ww = np.meshgrid(u, w, indexing="ij")
    vv = np.stack([(R + r * np.cos(ww)) * np.cos(uu), (R + r * np.cos(ww)) * np.sin(uu), r * np.sin(ww)], -1).reshape(-1, 3).astype(dtype)
    f = []
    for i in range(nu):
        for j in range(nv):
            a, b = i * nv + j, ((i + 1) % nu) * nv + j
            c, d = ((i + 1) % nu) * nv + (j + 1) % nv, i * nv + (j + 1) % nv
            f += [(a, b, c), (a, c, d)]
    return TriangleOctree(np.array(f, dtype=np.int64), vv, dtype=dtype)
