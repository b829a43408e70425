"""whatsthepoint.jl_b200 — B200 (sm_100a) implementation of the WhatsThePoint.jl hot path:
stencil connectivity (set_topology with KNNTopology / RadiusTopology) and node repulsion
(repel). See DESIGN.md. The directory name contains a dot, so load it through
`__graft_entry__.load_package()` (registered in sys.modules as `wtp_b200`).
"""
from . import _lib
from ._lib import Context, WtpArgumentError, WtpError, default_context, shard_range
from .api import (AbstractSpacing, AbstractTopology, BoundaryLayerSpacing, ClippedSpacingForce, ConstantSpacing, CSRRows,
                  FlatRows, InverseDistanceForce, KNNTopology, LogLike, NoTopology, PointBoundary, PointCloud,
                  PointSurface, PointVolume, RadiusTopology, RepelForceModel, SpacingEquilibriumForce,
                  StrongSpacingForce, compute_force, compute_normals, gradient_limit_field, hastopology, isinside, metrics, neighbors, points, rebuild_topology_, repel,
                  search, searchdists, set_topology, spacing_fidelity_metrics, spacing_metrics, topology)
from .mesh import TriangleOctree, cuboid_mesh, icosphere_mesh, read_binary_stl, torus_mesh, unit_cube_mesh

__all__ = [n for n in dir() if not n.startswith("_")]
