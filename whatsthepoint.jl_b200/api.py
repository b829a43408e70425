"""Host-side mirror of the reference interface for the hot path.

Julia is not available in this image, so the reference-facing layer above the C ABI is
written in Python with the reference's names, argument meaning and error behaviour:

    set_topology(cloud, KNNTopology, k)        src/cloud.jl:200, src/volume.jl:97, src/surface.jl:167
    set_topology(cloud, RadiusTopology, r)     src/cloud.jl:212, src/volume.jl:109, src/surface.jl:184
    rebuild_topology_(cloud)  (rebuild_topology!)   src/cloud.jl:224, src/topology.jl:109-129
    neighbors / hastopology / topology / points     src/cloud.jl:171-197,235
    repel(cloud, spacing; ...)                 src/repel.jl:56-95
    repel(cloud, spacing, octree; ...)         src/repel.jl:122-181 (mesh wall rule, _reconstruct_cloud :590-629)
    isinside(points, octree)                   src/octree/triangle_octree.jl:97-115
    metrics(cloud; k)                          src/metrics.jl:19-41
    compute_normals(points; k)                 src/normals.jl:9-44
    _gradient_limit_field (gradient_limit_field)   src/discretization/algorithms/octree.jl:677-717
    ConstantSpacing / LogLike / BoundaryLayerSpacing   src/discretization/spacings.jl
    InverseDistanceForce / SpacingEquilibriumForce / ClippedSpacingForce / StrongSpacingForce
                                               src/repel_forces.jl

Unitful quantities do not exist here: coordinates and lengths are plain float32/float64
numbers (what `ustrip` gives the reference at this boundary). Neighbour indices keep the
reference's convention: 1-based, global order = boundary surfaces in insertion order,
then volume (src/cloud.jl:235-237).

Everything numerical is done by libwtp_cuda.so; nothing here computes on the CPU.
The Julia shim a maintainer would add is in julia/WTPCuda.jl.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

from . import _lib
from ._lib import WtpArgumentError, WtpError, default_context

log = logging.getLogger("whatsthepoint_b200")


# ------------------------------------------------------------------ topology
class AbstractTopology:
    """src/topology.jl:7"""


class NoTopology(AbstractTopology):
    """src/topology.jl:14"""

    def __repr__(self):
        return "NoTopology()"


class FlatRows:
    """N x k neighbour table behaving like Vector{Vector{Int}} (the storage parameter S of
    KNNTopology{S}, src/topology.jl:25): len(), rows[i] -> int64 view of length k."""

    def __init__(self, table: np.ndarray):
        self.table = table

    def __len__(self):
        return self.table.shape[0]

    def __getitem__(self, i):
        return self.table[i]

    def __iter__(self):
        return iter(self.table)


class CSRRows:
    """Ragged neighbour lists over one CSR buffer (RadiusTopology storage)."""

    def __init__(self, offsets: np.ndarray, indices: np.ndarray):
        self.offsets, self.indices = offsets, indices

    def __len__(self):
        return self.offsets.shape[0] - 1

    def __getitem__(self, i):
        if i < 0:
            i += len(self)
        return self.indices[self.offsets[i]:self.offsets[i + 1]]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class KNNTopology(AbstractTopology):
    """src/topology.jl:25-28 (mutable: rebuild_topology! reassigns .neighbors)"""

    def __init__(self, neighbors, k: int):
        self.neighbors, self.k = neighbors, int(k)

    def __repr__(self):
        return f"KNNTopology(k={self.k})"

    def show(self) -> str:  # Base.show(io, MIME"text/plain", t), src/topology.jl:131-136
        return f"KNNTopology\n├─k: {self.k}\n└─points: {len(self.neighbors)}\n"


class RadiusTopology(AbstractTopology):
    """src/topology.jl:39-42"""

    def __init__(self, neighbors, radius):
        self.neighbors, self.radius = neighbors, radius

    def __repr__(self):
        return f"RadiusTopology(r={self.radius})"

    def show(self) -> str:
        return f"RadiusTopology\n├─radius: {self.radius}\n└─points: {len(self.neighbors)}\n"


def _get_radius(radius, pts):
    """src/topology.jl:99-100: a number, or a function of the point set returning one."""
    return radius(pts) if callable(radius) else radius


def _build_knn_neighbors(pts: np.ndarray, k: int, ctx=None) -> FlatRows:
    """src/topology.jl:79-84 -> wtp_knn_{f32,f64}."""
    ctx = ctx or default_context()
    return FlatRows(ctx.knn(pts, int(k)))


def _build_radius_neighbors(pts: np.ndarray, radius, ctx=None) -> CSRRows:
    """src/topology.jl:91-97 -> wtp_radius_count/fill."""
    ctx = ctx or default_context()
    r = float(_get_radius(radius, pts))
    return CSRRows(*ctx.radius(pts, r))


# ---------------------------------------------------------------- data model
def _coords(x, dtype=None) -> np.ndarray:
    a = np.asarray(x)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    if dtype is not None:
        a = a.astype(dtype, copy=False)
    if a.ndim != 2 or a.shape[1] not in (2, 3):
        raise WtpArgumentError(1, f"points must be N x 2 or N x 3, got {a.shape}")
    return np.ascontiguousarray(a)


class _HasTopology:
    topology: AbstractTopology

    def _points(self) -> np.ndarray:
        raise NotImplementedError

    def _with_topology(self, topo):
        raise NotImplementedError


@dataclass
class PointSurface(_HasTopology):
    """src/surface.jl:33-56: points (+ normals, areas) of one named surface."""
    points: np.ndarray
    normals: Optional[np.ndarray] = None
    areas: Optional[np.ndarray] = None
    topology: AbstractTopology = field(default_factory=NoTopology)

    def __post_init__(self):
        self.points = _coords(self.points)

    def __len__(self):
        return self.points.shape[0]

    def _points(self):
        return self.points

    def _with_topology(self, topo):
        return PointSurface(self.points, self.normals, self.areas, topo)


@dataclass
class PointVolume(_HasTopology):
    """src/volume.jl:10-20"""
    points: np.ndarray
    topology: AbstractTopology = field(default_factory=NoTopology)

    def __post_init__(self):
        self.points = _coords(self.points)

    def __len__(self):
        return self.points.shape[0]

    def _points(self):
        return self.points

    def _with_topology(self, topo):
        return PointVolume(self.points, topo)


class PointBoundary:
    """src/boundary.jl: ordered dict of named surfaces."""

    def __init__(self, surfaces):
        if isinstance(surfaces, dict):
            self.surfaces = {k: (v if isinstance(v, PointSurface) else PointSurface(v)) for k, v in surfaces.items()}
        elif isinstance(surfaces, PointSurface):
            self.surfaces = {"surface1": surfaces}
        else:
            self.surfaces = {"surface1": PointSurface(surfaces)}

    def __len__(self):
        return sum(len(s) for s in self.surfaces.values())

    def _points(self):  # mapreduce(points, vcat, surfaces), src/boundary.jl:164
        return np.concatenate([s.points for s in self.surfaces.values()], axis=0)


class PointCloud(_HasTopology):
    """src/cloud.jl:11-15: boundary + volume + topology; set_topology returns a new cloud
    sharing the point storage (:204)."""

    def __init__(self, boundary, volume=None, topology: AbstractTopology | None = None):
        self.boundary = boundary if isinstance(boundary, PointBoundary) else PointBoundary(boundary)
        d = self.boundary._points().shape[1]
        dt = self.boundary._points().dtype
        if volume is None:
            volume = PointVolume(np.zeros((0, d), dtype=dt))
        elif not isinstance(volume, PointVolume):
            volume = PointVolume(volume)
        # boundary and volume are promoted to one machine type at construction (src/cloud.jl:38-56)
        common = np.result_type(dt, volume.points.dtype)
        if volume.points.dtype != common:
            volume = PointVolume(volume.points.astype(common), volume.topology)
        if dt != common:
            self.boundary = PointBoundary({k: PointSurface(s.points.astype(common), s.normals, s.areas, s.topology)
                                           for k, s in self.boundary.surfaces.items()})
        self.volume = volume
        self.topology = topology if topology is not None else NoTopology()

    def __len__(self):
        return len(self.boundary) + len(self.volume)

    def _points(self):  # vcat(points(boundary), points(volume)), src/cloud.jl:235-237
        return np.concatenate([self.boundary._points(), self.volume.points], axis=0)

    def _with_topology(self, topo):
        return PointCloud(self.boundary, self.volume, topo)

    def show(self) -> str:
        return f"PointCloud\n├─{len(self)} points\n└─Topology: {type(self.topology).__name__}\n"


def points(x) -> np.ndarray:
    return x._points()


def topology(x) -> AbstractTopology:
    return x.topology


def hastopology(x) -> bool:
    return not isinstance(x.topology, NoTopology)


def neighbors(x, i: int | None = None):
    """src/topology.jl:52-62, src/cloud.jl:185-192. `i` is 1-based like the reference."""
    t = x if isinstance(x, AbstractTopology) else x.topology
    if isinstance(t, NoTopology):
        raise WtpArgumentError(1, "NoTopology has no neighbors")
    return t.neighbors if i is None else t.neighbors[i - 1]


def set_topology(x, kind, param, ctx=None):
    """set_topology(x, KNNTopology, k) / set_topology(x, RadiusTopology, radius) for
    PointCloud, PointVolume and PointSurface; returns a new container."""
    pts = x._points()
    if kind is KNNTopology:
        if not isinstance(param, (int, np.integer)):
            raise TypeError("set_topology(x, KNNTopology, k): k must be an Int")
        return x._with_topology(KNNTopology(_build_knn_neighbors(pts, int(param), ctx), int(param)))
    if kind is RadiusTopology:
        return x._with_topology(RadiusTopology(_build_radius_neighbors(pts, param, ctx), param))
    raise TypeError(f"unknown topology type {kind!r}")


def rebuild_topology_(x, ctx=None) -> None:
    """rebuild_topology!(x): re-run the builder with the stored parameter and assign into
    the (mutable) topology object (src/topology.jl:109-129)."""
    t = x.topology
    pts = x._points()
    if isinstance(t, KNNTopology):
        t.neighbors = _build_knn_neighbors(pts, t.k, ctx)
    elif isinstance(t, RadiusTopology):
        t.neighbors = _build_radius_neighbors(pts, t.radius, ctx)
    return None


def search(x, k: int, ctx=None) -> np.ndarray:
    """search(cloud, KNearestSearch(cloud, k)): k nearest INCLUDING self, 1-based (src/neighbors.jl:9-14)."""
    ctx = ctx or default_context()
    return ctx.knn(x._points() if hasattr(x, "_points") else x, int(k), include_self=True)


def searchdists(x, k: int, ctx=None):
    """searchdists(cloud, KNearestSearch(cloud, k)) -> (indices, distances) (src/neighbors.jl:16-21)."""
    ctx = ctx or default_context()
    return ctx.knn(x._points() if hasattr(x, "_points") else x, int(k), include_self=True, dists=True)


# ------------------------------------------------------------------ spacings
class AbstractSpacing:
    def _abi(self, dtype):
        raise NotImplementedError

    def __call__(self, pts, ctx=None):
        """spacing.(points): evaluated on the device (wtp_spacing_eval)."""
        ctx = ctx or default_context()
        pts = np.ascontiguousarray(pts)
        single = pts.ndim == 1
        p2 = pts[None, :] if single else pts
        sp, keep = self._abi(p2.dtype)
        out = ctx.spacing_eval(sp, p2)
        del keep
        return out[0] if single else out


class ConstantSpacing(AbstractSpacing):
    """src/discretization/spacings.jl:35-39"""

    def __init__(self, dx):
        self.dx = dx

    def _abi(self, dtype):
        return _lib.Context.make_spacing("constant", a=self.dx)


class LogLike(AbstractSpacing):
    """src/discretization/spacings.jl:49-72: h = h0 * d / (a + d), a = h0 * (2 - g)."""

    def __init__(self, boundary_points, base_size, growth_rate):
        bp = points(boundary_points) if hasattr(boundary_points, "_points") else np.asarray(boundary_points)
        if len(bp) == 0:
            raise WtpArgumentError(1, "boundary_points must be non-empty")
        self.boundary, self.base_size, self.growth_rate = _coords(bp), base_size, growth_rate

    def _abi(self, dtype):
        return _lib.Context.make_spacing("loglike", a=self.base_size, b=self.growth_rate, bnd_pts=self.boundary.astype(dtype, copy=False))


class BoundaryLayerSpacing(AbstractSpacing):
    """src/discretization/spacings.jl:93-133"""

    def __init__(self, boundary_points, *, at_wall, bulk, layer_thickness):
        bp = points(boundary_points) if hasattr(boundary_points, "_points") else np.asarray(boundary_points)
        if len(bp) == 0:
            raise WtpArgumentError(1, "boundary_points must be non-empty")
        if not layer_thickness > 0:
            raise WtpArgumentError(1, f"layer_thickness must be positive, got {layer_thickness}")
        self.boundary, self.at_wall, self.bulk, self.layer_thickness = _coords(bp), at_wall, bulk, layer_thickness

    def _abi(self, dtype):
        return _lib.Context.make_spacing("boundary_layer", a=self.at_wall, b=self.bulk, c=self.layer_thickness,
                                         bnd_pts=self.boundary.astype(dtype, copy=False))


# -------------------------------------------------------------------- forces
class RepelForceModel:
    """src/repel_forces.jl:9"""
    kind: str = ""
    beta: float = 0.2
    u0: float = 1.0
    gamma: float = 3.0

    def _abi(self):
        return _lib.Context.make_force(self.kind, self.beta, self.u0, self.gamma)


class InverseDistanceForce(RepelForceModel):
    kind = "inverse"

    def __init__(self, beta=0.2):
        self.beta = beta


class SpacingEquilibriumForce(RepelForceModel):
    kind = "equilibrium"

    def __init__(self, beta=0.2):
        self.beta = beta


class ClippedSpacingForce(RepelForceModel):
    kind = "clipped"

    def __init__(self, beta=0.2, u0=1.0):
        self.beta, self.u0 = beta, u0


class StrongSpacingForce(RepelForceModel):
    kind = "strong"

    def __init__(self, beta=0.2, gamma=3.0):
        self.beta, self.gamma = beta, gamma


def compute_force(model: RepelForceModel, u, ctx=None):
    """compute_force(model, u) on the device (src/repel_forces.jl:22)."""
    if not isinstance(model, RepelForceModel) or model.kind not in _lib.FORCE_KINDS:
        raise WtpError(3, "user-defined RepelForceModel subtypes cannot cross the C ABI (no CPU fallback)")
    ctx = ctx or default_context()
    arr = np.atleast_1d(np.asarray(u))
    if arr.dtype not in (np.float32, np.float64):
        arr = arr.astype(np.float64)
    out = ctx.force_eval(model._abi(), arr)
    return out[0] if np.ndim(u) == 0 else out


# --------------------------------------------------------------------- repel
def repel(cloud: PointCloud, spacing: AbstractSpacing, octree=None, *, beta=0.2, force_model: RepelForceModel | None = None,
          alpha=None, alpha_min=None, k=21, max_iters=1000, tol=1.0e-6, rebuild_every: int = 1, cull_ratio=0.0,
          kick_after: int = 0, stall_after: int = 50, cv_target=0.0, deposit_ratio=0.0, convergence: list | None = None,
          trace: list | None = None, isinside: Callable | bool | None = None, kick_seed: int = 0, ctx=None) -> PointCloud:
    """repel(cloud, spacing; kwargs...) (src/repel.jl:56-95): volume points move, boundary
    points are the fixed wall; returns a new cloud with NoTopology.

    The relaxation (_relax!, src/repel.jl:202-339) runs in libwtp_cuda.so, and so does the survivor
    filter `filter(x -> isinside(x, cloud), p)` (:90), which the reference always applies: Green's
    function over the boundary elements in 3-D (it needs their normals and areas, which every
    reference PointSurface carries), winding number in 2-D. `isinside=False` opts out (a mirror-only
    convenience for clouds built from bare coordinates); a callable is used as the predicate on the
    N x D array of moved points. `cull_ratio > 0` applies the near-duplicate cull (:91-93, :549-580).
    """
    if rebuild_every < 1:
        raise WtpArgumentError(1, "rebuild_every must be ≥ 1")                     # src/repel.jl:74
    if octree is not None and not hasattr(octree, "feature_normals"):
        raise WtpError(3, "repel(cloud, spacing, octree): octree must be a TriangleOctree (only its TriangleIndex arrays cross the C ABI)")
    if deposit_ratio < 0:
        raise WtpArgumentError(1, "deposit_ratio must be ≥ 0")                     # src/repel.jl:143
    if deposit_ratio > 0 and octree is None:
        raise TypeError("deposit_ratio belongs to repel(cloud, spacing, octree)")      # keyword of the 3-argument method only, :138
    if not isinstance(spacing, AbstractSpacing):
        raise WtpError(3, "only ConstantSpacing, LogLike and BoundaryLayerSpacing can cross the C ABI (no CPU fallback)")
    ctx = ctx or default_context()
    fm = force_model if force_model is not None else ClippedSpacingForce(beta)    # :60
    if not isinstance(fm, RepelForceModel) or fm.kind not in _lib.FORCE_KINDS:
        raise WtpError(3, "user-defined RepelForceModel subtypes cannot cross the C ABI (no CPU fallback)")
    bnd_p = cloud.boundary._points()
    n_bnd = bnd_p.shape[0]
    snap = np.concatenate([bnd_p, cloud.volume.points], axis=0)                    # :80 / :146-148
    dtype = snap.dtype
    if octree is not None and snap.shape[1] != 3:
        raise TypeError("repel(cloud, spacing, octree) is defined for 3-D clouds only")   # dispatch on 𝔼{3}, :123
    sp, keep = spacing._abi(dtype)
    if alpha is None:                                                              # :61  α = minimum(spacing.(to(cloud)))/20
        alpha = dtype.type(ctx.spacing_eval(sp, snap).min()) / dtype.type(20)
    if alpha_min is None:                                                          # :62
        alpha_min = alpha / 100
    is_bnd = None
    if octree is not None:                                                         # every point moves, :172; is_bnd :155
        is_bnd = np.arange(snap.shape[0]) < n_bnd
    new_snap, conv, res, tr = ctx.repel(snap, 0 if octree is not None else n_bnd, sp, fm._abi(), k=k, max_iters=max_iters, tol=tol,
                                        rebuild_every=rebuild_every, stall_after=stall_after, cv_target=cv_target,
                                        alpha_lo=alpha_min, alpha_max=alpha, kick_after=kick_after, trace=trace is not None,
                                        mesh=octree, is_bnd=is_bnd, n_protected=n_bnd, kick_seed=kick_seed,
                                        deposit_ratio=deposit_ratio if octree is not None else 0.0)
    del keep
    if convergence is not None:
        convergence.extend(float(c) for c in conv)                                 # :88
    if trace is not None:
        trace.extend(tr)
    i = res["iters"]
    if res["stop_reason"] == "cv_target":                                          # the reference's @info/@warn lines
        log.info("Node repel stopped in %d iterations: spacing CV target reached", i)
    elif res["stop_reason"] == "stall":
        log.info("Node repel stopped in %d iterations: spacing CV stalled for %d iterations", i, stall_after)
    elif res["stop_reason"] == "tol":
        log.info("Node repel finished in %d iterations", i)
    elif max_iters > 0:
        log.warning("Node repel reached maximum iterations")
    if octree is not None:
        keep_mask = _cull(new_snap, spacing, cull_ratio, ctx) if cull_ratio > 0 else None   # :179
        is_bnd = ctx.last_wall["is_bnd"].astype(bool)                              # deposition converts volume points (:152-155, :509)
        out = _reconstruct_cloud(cloud, new_snap, ctx.last_wall["tri_indices"], is_bnd, n_bnd, octree, keep_mask, spacing, ctx)
        out.repel_result = res
        out.escaped = ctx.last_wall["escaped"].astype(bool)
        return out
    p = new_snap[n_bnd:]
    if isinside is not False and len(p) > 0:                                       # :90  filter(x -> isinside(x, cloud), p)
        keep_mask = isinside(p) if callable(isinside) else globals()["isinside"](p, cloud, ctx=ctx)
        p = p[np.asarray(keep_mask, dtype=bool)]
    if cull_ratio > 0 and len(p) > 0:                                              # :91-93
        p = p[_cull(p, spacing, cull_ratio, ctx)]
    out = PointCloud(cloud.boundary, PointVolume(p), NoTopology())                 # :94
    out.repel_result = res
    return out


def _cull(pts: np.ndarray, spacing: AbstractSpacing, ratio, ctx) -> np.ndarray:
    """_cull(pts, spacing, ratio) (src/repel.jl:549-556): near-duplicate keep mask plus the defect warning."""
    sp, keep = spacing._abi(pts.dtype)
    s = ctx.spacing_eval(sp, pts)
    del keep
    mask = ctx.cull_mask(pts, s, float(ratio))
    n_culled = int((~mask).sum())
    if n_culled > 0:
        log.warning("Cull removed %d near-duplicate point(s) — repel left defects behind (cull_ratio = %s)", n_culled, ratio)
    return mask


def _reconstruct_cloud(cloud: PointCloud, p: np.ndarray, tri_indices: np.ndarray, is_bnd: np.ndarray, n_boundary: int, octree,
                       keep: np.ndarray | None = None, spacing=None, ctx=None) -> PointCloud:
    """src/repel.jl:590-629: kept points are split by is_bnd into one `boundary` surface and the
    volume; projected boundary points take the landing triangle's normal, imported ones keep
    their area, deposited ones (id > n_boundary) get spacing²."""
    if keep is None:
        keep = np.ones(len(p), dtype=bool)
    normals = [s.normals for s in cloud.boundary.surfaces.values()]
    areas = [s.areas for s in cloud.boundary.surfaces.values()]
    orig_normals = np.concatenate(normals, axis=0) if all(x is not None for x in normals) and normals else None
    orig_areas = np.concatenate(areas, axis=0) if all(x is not None for x in areas) and areas else None
    b = np.flatnonzero(is_bnd & keep)
    tri = tri_indices[b]
    face = octree.astype(p.dtype).face
    new_normals = face[np.maximum(tri, 1) - 1].copy()
    if orig_normals is not None:
        keep_orig = (tri == 0) & (b < n_boundary)
        new_normals[keep_orig] = orig_normals[b[keep_orig]]
    new_areas = None
    if orig_areas is not None:
        new_areas = np.empty(len(b), dtype=orig_areas.dtype)
        imported = b < n_boundary
        new_areas[imported] = orig_areas[b[imported]]
        if (~imported).any():                                                      # deposited points: spacing(p[id])^2 (:617)
            sp, keepalive = spacing._abi(p.dtype)
            new_areas[~imported] = (ctx or default_context()).spacing_eval(sp, p[b[~imported]]) ** 2
            del keepalive
    surf = PointSurface(p[b], new_normals, new_areas)
    return PointCloud(PointBoundary({"boundary": surf}), PointVolume(p[~is_bnd & keep]), NoTopology())


def _boundary_elements(domain):
    """(points, normals, areas) of all surfaces of a PointCloud / PointBoundary / PointSurface, concatenated."""
    if isinstance(domain, PointCloud):
        domain = domain.boundary
    surfaces = list(domain.surfaces.values()) if isinstance(domain, PointBoundary) else [domain]
    pts = np.concatenate([s.points for s in surfaces], axis=0)
    if pts.shape[1] == 2:
        return pts, None, None
    if any(s.normals is None or s.areas is None for s in surfaces):
        raise WtpArgumentError(1, "the 3-D isinside (Green's function, src/isinside.jl:86-106) needs the normal and area of every "
                                  "boundary element; repel(cloud, spacing) applies it to the moved points (src/repel.jl:90) — "
                                  "build the surfaces with normals and areas, or pass isinside=False")
    return pts, np.concatenate([s.normals for s in surfaces], axis=0), np.concatenate([s.areas for s in surfaces], axis=0)


def isinside(pts, domain, ctx=None) -> np.ndarray:
    """isinside(points, octree) (src/octree/triangle_octree.jl:97-115) or isinside(points, cloud | boundary | surface)
    (src/isinside.jl: Green's function over the boundary elements in 3-D, winding number in 2-D), on the device."""
    ctx = ctx or default_context()
    a = pts._points() if hasattr(pts, "_points") else np.asarray(pts)
    single = a.ndim == 1
    q = a[None, :] if single else a
    if hasattr(domain, "feature_normals"):
        out = ctx.mesh_isinside(domain, q)
    else:
        bx, bn, ba = _boundary_elements(domain)
        out = ctx.isinside(q.astype(bx.dtype, copy=False), bx, bn, ba)
    return bool(out[0]) if single else out


# ------------------------------------------------------- other consumers of the index
def compute_normals(x, k: int = 5, ctx=None) -> np.ndarray:
    """compute_normals(surf | points; k) (src/normals.jl:9-44): unit PCA normals (Hoppe 1992) from the covariance of each
    point's k nearest points, itself included; unoriented (orient_normals!, :75-117, is a serial graph walk on the host)."""
    ctx = ctx or default_context()
    pts = x._points() if hasattr(x, "_points") else _coords(x)
    return ctx.normals(pts, int(k))


def gradient_limit_field(centers, h0, g, k: int = 12, tol: float = 1.0e-3, max_sweeps: int = 2000, ctx=None) -> np.ndarray:
    """_gradient_limit_field (src/discretization/algorithms/octree.jl:677-717) on the leaf centres: the g-Lipschitz
    envelope of h0 by min-plus sweeps over the k-NN graph of the centres."""
    ctx = ctx or default_context()
    h, _ = ctx.gradient_limit(_coords(centers), h0, float(g), int(k), float(tol), int(max_sweeps))
    return h


# ------------------------------------------------------------------- metrics
def metrics(cloud, k: int = 20, ctx=None, verbose: bool = True) -> dict:
    """metrics(cloud; k) (src/metrics.jl:19-41)."""
    ctx = ctx or default_context()
    m = ctx.metrics(cloud._points() if hasattr(cloud, "_points") else cloud, int(k))
    m["k"] = int(k)
    if verbose:
        print("Cloud Metrics\n-------------")
        print(f"avg. distance to {k} nearest neighbors: {m['avg']}")
        print(f"std. distance to {k} nearest neighbors: {m['std']}")
        print(f"max. distance to {k} nearest neighbors: {m['max']}")
        print(f"min. distance to {k} nearest neighbors: {m['min']}")
        print(f"separation (min nearest-neighbor distance): {m['separation']}")
        print(f"fill (max nearest-neighbor distance):       {m['fill']}")
        print(f"mesh ratio (fill / separation, ≥1):         {m['mesh_ratio']}")
    return m


def spacing_metrics(cloud, spacing: AbstractSpacing, k: int = 20, ctx=None) -> dict:
    """spacing_metrics(cloud, spacing; k) (src/metrics.jl:56-71) -> max_error, mean_error, std_error, k."""
    if not isinstance(spacing, AbstractSpacing):
        raise WtpError(3, "only ConstantSpacing, LogLike and BoundaryLayerSpacing can cross the C ABI (no CPU fallback)")
    ctx = ctx or default_context()
    pts = cloud._points() if hasattr(cloud, "_points") else _coords(cloud)
    sp, keep = spacing._abi(pts.dtype)
    m = ctx.spacing_metrics(pts, sp, int(k))
    del keep
    m["k"] = int(k)
    return m


def spacing_fidelity_metrics(cloud, spacing: AbstractSpacing, k: int = 30, coord_radius: float = 1.4, ctx=None) -> dict:
    """spacing_fidelity_metrics(cloud, spacing; k, coord_radius) (src/metrics.jl:88-129) -> mean_dnn_h, cv, p05, p50, p95,
    coordination, k, coord_radius."""
    if not isinstance(spacing, AbstractSpacing):
        raise WtpError(3, "only ConstantSpacing, LogLike and BoundaryLayerSpacing can cross the C ABI (no CPU fallback)")
    ctx = ctx or default_context()
    pts = cloud._points() if hasattr(cloud, "_points") else _coords(cloud)
    sp, keep = spacing._abi(pts.dtype)
    m = ctx.spacing_fidelity_metrics(pts, sp, int(k), float(coord_radius))
    del keep
    m["k"], m["coord_radius"] = min(int(k), len(pts)), coord_radius
    return m
