"""ctypes binding of libwtp_cuda.so (include/wtp_cuda.h).

There is no CPU fallback: if the shared library is missing or no B200 is usable the
import of the library / creation of a context raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwtp_cuda.so")

WTP_MAX_K = 256

STATUS = {0: "ok", 1: "bad_arg", 2: "k_too_large", 3: "unsupported", 4: "cuda", 5: "nccl", 6: "oom", 7: "state"}
FORCE_KINDS = {"inverse": 0, "equilibrium": 1, "clipped": 2, "strong": 3}
SPACING_KINDS = {"constant": 0, "loglike": 1, "boundary_layer": 2}
STOP_REASONS = {0: "max_iters", 1: "tol", 2: "cv_target", 3: "stall"}


class WtpError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"libwtp_cuda: {STATUS.get(status, status)}: {message}")
        self.status = status


class WtpArgumentError(WtpError, ValueError):
    """Maps to the reference's ArgumentError (src/repel.jl:74,143, src/topology.jl:61-62)."""


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


class SpacingMetrics(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("max_error", "mean_error", "std_error")]


class SpacingFidelity(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("mean_dnn_h", "cv", "p05", "p50", "p95", "coordination")]


class CloudMetrics(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("avg", "std", "max", "min", "separation", "fill", "mesh_ratio")]


class Timing(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("ms_h2d", "ms_bbox", "ms_cellkey", "ms_sort", "ms_reorder", "ms_query",
                                         "ms_scan", "ms_reduce", "ms_comm", "ms_d2h", "ms_total")] + \
               [("sort_passes", C.c_int32), ("query_launches", C.c_int32), ("n_cells", C.c_int64),
                ("n_ring_expanded", C.c_int64), ("n_leftover_sparse", C.c_int64), ("n_leftover_dense", C.c_int64),
                ("n_leftover_other", C.c_int64), ("n_window_points", C.c_int64), ("n_window_missed", C.c_int64),
                ("n_peer_ranks", C.c_int64), ("bytes_h2d", C.c_int64), ("bytes_d2h", C.c_int64)]


_lib = None


def load() -> C.CDLL:
    """Load libwtp_cuda.so; raises (no fallback) if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a). "
                              "There is no CPU fallback for the WhatsThePoint hot path.")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        lib.wtp_last_error.restype = C.c_char_p
        lib.wtp_status_string.restype = C.c_char_p
        lib.wtp_launch_count.restype = C.c_int64
        lib.wtp_radius_nnz.restype = C.c_int64
        lib.wtp_shard_begin.restype = C.c_int64
        lib.wtp_shard_end.restype = C.c_int64
        lib.wtp_shard_owned_count.restype = C.c_int64
        lib.wtp_shard_owned_count.argtypes = [C.c_void_p]
        lib.wtp_shard_begin.argtypes = [C.c_int64, C.c_int32, C.c_int32]
        lib.wtp_shard_end.argtypes = [C.c_int64, C.c_int32, C.c_int32]
        _lib = lib
    return _lib


def shard_range(n: int, rank: int, world: int):
    """Contiguous block partition used by every sharded entry point (pure arithmetic)."""
    if world <= 1:
        return 0, n
    return (n * rank) // world, (n * (rank + 1)) // world


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"point coordinates must be float32 or float64, got {dtype}")


def _as_points(pts) -> np.ndarray:
    pts = np.ascontiguousarray(pts)
    if pts.ndim != 2 or pts.shape[1] not in (2, 3):
        raise WtpArgumentError(1, f"points must be N x 2 or N x 3, got shape {pts.shape}")
    _sfx(pts.dtype)
    return pts


def _vp(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context:
    """One wtp_ctx: a CUDA device, a stream and the cached device buffers."""

    def __init__(self, device: int = 0, devices=None):
        """One device (wtp_create), or `devices=[0, 1, ...]`: one context over several GPUs of the box in this one process
        (wtp_create_multi) — the host entry points that shard are then answered by all of them together."""
        self._lib = load()
        self._h = C.c_void_p()
        if devices is not None and len(devices) > 1:
            arr = (C.c_int32 * len(devices))(*[int(d) for d in devices])
            rc = self._lib.wtp_create_multi(C.byref(self._h), arr, C.c_int32(len(devices)))
            if rc != 0:
                raise WtpError(rc, f"wtp_create_multi(devices={list(devices)}) failed: the devices must be usable sm_100 GPUs that can "
                                   "access each other's memory, with libnccl.so.2 loadable")
            device = int(devices[0])
        else:
            if devices is not None:
                device = int(devices[0])
            rc = self._lib.wtp_create(C.byref(self._h), C.c_int32(device))
            if rc != 0:
                raise WtpError(rc, f"wtp_create(device={device}) failed: no usable sm_100 GPU (there is no CPU fallback)")
        self.device = device
        self.devices = list(devices) if devices is not None else [device]
        self.rank, self.world = 0, 1

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.wtp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc != 0:
            msg = self._lib.wtp_last_error(self._h).decode()
            raise (WtpArgumentError if rc in (1, 2) else WtpError)(rc, msg)

    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self._lib.wtp_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def set_timing(self, enable: bool):
        self._check(self._lib.wtp_set_timing(self._h, C.c_int32(1 if enable else 0)))

    def timing(self) -> dict:
        t = Timing()
        self._check(self._lib.wtp_get_timing(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in Timing._fields_}

    def launch_count(self) -> int:
        return int(self._lib.wtp_launch_count(self._h))

    def set_cell_occupancy(self, m: float):
        self._check(self._lib.wtp_set_cell_occupancy(self._h, C.c_double(m)))

    def host_register(self, arr: np.ndarray):
        self._check(self._lib.wtp_host_register(self._h, C.c_void_p(arr.ctypes.data), C.c_int64(arr.nbytes)))

    def host_unregister(self, arr: np.ndarray):
        self._check(self._lib.wtp_host_unregister(self._h, C.c_void_p(arr.ctypes.data)))

    def comm_init(self, rank: int, world: int, unique_id: bytes | None):
        buf = (C.c_char * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        self._check(self._lib.wtp_comm_init(self._h, C.c_int32(rank), C.c_int32(world), buf))
        self.rank, self.world = rank, world

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_char * 128)()
        rc = load().wtp_comm_unique_id(buf)
        if rc != 0:
            raise WtpError(rc, "wtp_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return bytes(buf)

    def shard(self, n: int):
        return shard_range(n, self.rank, self.world)

    def owned(self) -> np.ndarray:
        """Caller indices (1-based) of the rows the last k-NN call of a sharded context answered, in the
        order of the compact device table (a contiguous run of the spatially sorted order)."""
        n = int(self._lib.wtp_shard_owned_count(self._h))
        ids = np.empty(max(n, 0), dtype=np.int64)
        self._check(self._lib.wtp_shard_owned(self._h, _vp(ids)))
        return ids

    # ------------------------------------------------------------ topology
    def knn(self, pts, k: int, *, dists: bool = False, include_self: bool = False, out_idx=None, out_dist=None):
        """_build_knn_neighbors (default) or search/searchdists (include_self). 1-based int64 N x k.
        On a sharded context only the rows of the points this rank owns (`owned()`) are filled."""
        pts = _as_points(pts)
        n, d = pts.shape
        idx = out_idx if out_idx is not None else np.empty((n, k), dtype=np.int64)
        dist = out_dist if out_dist is not None else (np.empty((n, k), dtype=pts.dtype) if dists else None)
        fn = getattr(self._lib, ("wtp_knn_self_" if include_self else "wtp_knn_") + _sfx(pts.dtype))
        self._check(fn(self._h, _vp(pts), C.c_int64(n), C.c_int32(d), C.c_int32(k), _vp(idx), _vp(dist)))
        return (idx, dist) if (dists or out_dist is not None) else idx

    def knn_dev(self, d_pts_ptr: int, n: int, d: int, k: int, dtype, d_out_idx_ptr: int, d_out_dist_ptr: int = 0, *, idx32: bool = False):
        """Device pointers in and out. idx32: the table is int32 (wtp_knn_dev_i32_*), int64 otherwise."""
        fn = getattr(self._lib, ("wtp_knn_dev_i32_" if idx32 else "wtp_knn_dev_") + _sfx(dtype))
        self._check(fn(self._h, C.c_void_p(d_pts_ptr), C.c_int64(n), C.c_int32(d), C.c_int32(k),
                       C.c_void_p(d_out_idx_ptr), C.c_void_p(d_out_dist_ptr or 0)))

    def radius(self, pts, r: float):
        """_build_radius_neighbors as CSR: (offsets int64[N+1] 0-based, indices int64[nnz] 1-based)."""
        pts = _as_points(pts)
        n, d = pts.shape
        sfx = _sfx(pts.dtype)
        b, e = self.shard(n)
        offsets = np.zeros(e - b + 1, dtype=np.int64)
        rr = C.c_float(r) if sfx == "f32" else C.c_double(r)
        self._check(getattr(self._lib, "wtp_radius_count_" + sfx)(self._h, _vp(pts), C.c_int64(n), C.c_int32(d), rr, _vp(offsets)))
        indices = np.empty(int(offsets[-1]), dtype=np.int64)
        self._check(self._lib.wtp_radius_fill(self._h, _vp(indices) if indices.size else None))
        return offsets, indices

    def radius_dev(self, d_pts_ptr: int, n: int, d: int, r: float, dtype, d_offsets_ptr: int):
        sfx = _sfx(dtype)
        rr = C.c_float(r) if sfx == "f32" else C.c_double(r)
        self._check(getattr(self._lib, "wtp_radius_count_dev_" + sfx)(self._h, C.c_void_p(d_pts_ptr), C.c_int64(n), C.c_int32(d), rr,
                                                                    C.c_void_p(d_offsets_ptr)))
        return int(self._lib.wtp_radius_nnz(self._h))

    def radius_fill_dev(self, d_indices_ptr: int):
        self._check(self._lib.wtp_radius_fill_dev(self._h, C.c_void_p(d_indices_ptr)))

    # --------------------------------------------------------------- repel
    @staticmethod
    def make_force(kind="clipped", beta=0.2, u0=1.0, gamma=3.0) -> Force:
        return Force(FORCE_KINDS[kind], float(beta), float(u0), float(gamma))

    @staticmethod
    def make_spacing(kind="constant", a=0.0, b=0.0, c=0.0, bnd_pts=None, bnd_ptr: int | None = None, n_bnd: int = 0):
        """Returns (Spacing, keepalive). bnd_pts: host array; bnd_ptr: device pointer (for *_dev calls)."""
        if bnd_pts is not None:
            bnd_pts = np.ascontiguousarray(bnd_pts)
            return Spacing(SPACING_KINDS[kind], float(a), float(b), float(c), bnd_pts.ctypes.data, bnd_pts.shape[0]), bnd_pts
        return Spacing(SPACING_KINDS[kind], float(a), float(b), float(c), bnd_ptr, n_bnd), None

    def repel(self, snap, n_fixed: int, sp: Spacing, force: Force, *, k=21, max_iters=1000, tol=1e-6, rebuild_every=1,
              stall_after=50, cv_target=0.0, alpha_lo, alpha_max, kick_after=0, trace=False, mesh=None, is_bnd=None,
              n_protected=None, kick_seed=0, deposit_ratio=0.0):
        """_relax! on snap = [fixed head; movable tail] (host array, copied). Returns
        (new_snap, conv, result dict, trace list | None)."""
        snap = np.array(_as_points(snap), copy=True)
        n_all, d = snap.shape
        n_move = n_all - n_fixed
        prm = RepelParams(k, max_iters, rebuild_every, stall_after, kick_after, 1 if mesh is not None else 0, 1 if trace else 0, 0,
                          float(alpha_lo), float(alpha_max), float(tol), float(cv_target),
                          int(n_fixed if n_protected is None else n_protected), int(kick_seed), float(deposit_ratio))
        conv = np.zeros(max(max_iters, 1), dtype=snap.dtype)
        tr = (TraceEntry * max(max_iters, 1))() if trace else None
        res = RepelResult()
        wall, keep = None, None
        self.last_wall = None
        if mesh is not None:   # repel(cloud, spacing, octree): boundary flags in, landing triangles / escape flags out
            flags = np.ascontiguousarray(is_bnd, dtype=np.uint8)
            tri_idx, esc = np.zeros(n_move, dtype=np.int64), np.zeros(n_move, dtype=np.uint8)
            w, keep = mesh.astype(snap.dtype).wall(flags, tri_idx, esc)
            wall = C.byref(w)
            self.last_wall = dict(tri_indices=tri_idx, escaped=esc, is_bnd=flags)   # is_bnd: rewritten by deposition
        fn = getattr(self._lib, "wtp_repel_" + _sfx(snap.dtype))
        self._check(fn(self._h, _vp(snap), C.c_int64(n_fixed), C.c_int64(n_move), C.c_int32(d), C.byref(sp), C.byref(force),
                       C.byref(prm), wall, _vp(conv), tr, C.byref(res)))
        del keep
        out_tr = None
        if trace:
            out_tr = [dict(iteration=i + 1, r=tr[i].r, s=tr[i].s, r_over_s=tr[i].r_over_s, idx_a=tr[i].idx_a, idx_b=tr[i].idx_b)
                      for i in range(res.iters)]
        return snap, conv[:res.iters].copy(), dict(iters=res.iters, stop_reason=STOP_REASONS[res.stop_reason],
                                                   last_cv=res.last_cv), out_tr

    def repel_dev(self, d_snap_ptr: int, n_fixed: int, n_move: int, d: int, dtype, sp: Spacing, force: Force, *, k=21,
                  max_iters=1000, tol=1e-6, rebuild_every=1, stall_after=50, cv_target=0.0, alpha_lo, alpha_max):
        prm = RepelParams(k, max_iters, rebuild_every, stall_after, 0, 0, 0, 0, float(alpha_lo), float(alpha_max),
                          float(tol), float(cv_target))
        conv = np.zeros(max(max_iters, 1), dtype=dtype)
        res = RepelResult()
        fn = getattr(self._lib, "wtp_repel_dev_" + _sfx(dtype))
        self._check(fn(self._h, C.c_void_p(d_snap_ptr), C.c_int64(n_fixed), C.c_int64(n_move), C.c_int32(d), C.byref(sp),
                       C.byref(force), C.byref(prm), _vp(conv), None, C.byref(res)))
        return conv[:res.iters].copy(), dict(iters=res.iters, stop_reason=STOP_REASONS[res.stop_reason], last_cv=res.last_cv)

    def spacing_eval(self, sp: Spacing, pts):
        pts = _as_points(pts)
        out = np.empty(pts.shape[0], dtype=pts.dtype)
        self._check(getattr(self._lib, "wtp_spacing_eval_" + _sfx(pts.dtype))(
            self._h, C.byref(sp), _vp(pts), C.c_int64(pts.shape[0]), C.c_int32(pts.shape[1]), _vp(out)))
        return out

    def force_eval(self, force: Force, u):
        u = np.ascontiguousarray(u)
        out = np.empty_like(u)
        self._check(getattr(self._lib, "wtp_force_eval_" + _sfx(u.dtype))(self._h, C.byref(force), _vp(u), C.c_int64(u.size), _vp(out)))
        return out

    def isinside(self, pts, bnd_pts, bnd_normals=None, bnd_areas=None, *, sums=False):
        """isinside(points, cloud) (src/isinside.jl): Green's function over the boundary elements in 3-D, winding
        number over the ordered boundary polygon in 2-D. Returns flags (and the raw sums)."""
        pts = _as_points(pts)
        d = pts.shape[1]
        bx = np.ascontiguousarray(bnd_pts, dtype=pts.dtype)
        bn = np.ascontiguousarray(bnd_normals, dtype=pts.dtype) if bnd_normals is not None else None
        ba = np.ascontiguousarray(bnd_areas, dtype=pts.dtype) if bnd_areas is not None else None
        out = np.zeros(pts.shape[0], dtype=np.uint8)
        s = np.zeros(pts.shape[0], dtype=pts.dtype) if sums else None
        fn = getattr(self._lib, "wtp_isinside_" + _sfx(pts.dtype))
        self._check(fn(self._h, _vp(pts), C.c_int64(pts.shape[0]), C.c_int32(d), _vp(bx), _vp(bn), _vp(ba), C.c_int64(bx.shape[0]), _vp(out), _vp(s)))
        return (out.astype(bool), s) if sums else out.astype(bool)

    def mesh_isinside(self, mesh, pts) -> np.ndarray:
        """isinside(points, octree) -> bool array (src/octree/triangle_octree.jl:97-115)."""
        pts = _as_points(pts)
        w, keep = mesh.astype(pts.dtype).wall()
        out = np.zeros(pts.shape[0], dtype=np.uint8)
        self._check(getattr(self._lib, "wtp_mesh_isinside_" + _sfx(pts.dtype))(self._h, C.byref(w), _vp(pts), C.c_int64(pts.shape[0]), _vp(out)))
        return out.astype(bool)

    def mesh_project(self, mesh, pts):
        """_project_to_boundary for a batch -> (projected points, 1-based triangle ids)."""
        pts = _as_points(pts)
        w, keep = mesh.astype(pts.dtype).wall()
        out, tri = np.empty_like(pts), np.zeros(pts.shape[0], dtype=np.int64)
        self._check(getattr(self._lib, "wtp_mesh_project_" + _sfx(pts.dtype))(self._h, C.byref(w), _vp(pts), C.c_int64(pts.shape[0]), _vp(out), _vp(tri)))
        return out, tri

    def cull_mask(self, pts, spacings, ratio: float) -> np.ndarray:
        """_near_duplicate_keep_mask(pts, spacings, ratio) (src/repel.jl:565-580) -> bool keep mask."""
        pts = _as_points(pts)
        sp = np.ascontiguousarray(spacings, dtype=pts.dtype)
        keep = np.ones(pts.shape[0], dtype=np.uint8)
        fn = getattr(self._lib, "wtp_cull_mask_" + _sfx(pts.dtype))
        self._check(fn(self._h, _vp(pts), C.c_int64(pts.shape[0]), C.c_int32(pts.shape[1]), _vp(sp), C.c_double(ratio), _vp(keep)))
        return keep.astype(bool)

    def spacing_metrics(self, pts, sp: Spacing, k=20) -> dict:
        """spacing_metrics(cloud, spacing; k) (src/metrics.jl:56-71)."""
        pts = _as_points(pts)
        out = SpacingMetrics()
        fn = getattr(self._lib, "wtp_spacing_metrics_" + _sfx(pts.dtype))
        self._check(fn(self._h, _vp(pts), C.c_int64(pts.shape[0]), C.c_int32(pts.shape[1]), C.c_int32(k), C.byref(sp), C.byref(out)))
        return {n: getattr(out, n) for n, _ in SpacingMetrics._fields_}

    def spacing_fidelity_metrics(self, pts, sp: Spacing, k=30, coord_radius=1.4) -> dict:
        """spacing_fidelity_metrics(cloud, spacing; k, coord_radius) (src/metrics.jl:88-129)."""
        pts = _as_points(pts)
        out = SpacingFidelity()
        fn = getattr(self._lib, "wtp_spacing_fidelity_" + _sfx(pts.dtype))
        self._check(fn(self._h, _vp(pts), C.c_int64(pts.shape[0]), C.c_int32(pts.shape[1]), C.c_int32(k), C.c_double(coord_radius),
                       C.byref(sp), C.byref(out)))
        return {n: getattr(out, n) for n, _ in SpacingFidelity._fields_}

    def normals(self, pts, k: int = 5) -> np.ndarray:
        """compute_normals(points; k) (src/normals.jl:9-44): unit PCA normals, N x D (unoriented: first nonzero component positive)."""
        pts = _as_points(pts)
        out = np.empty_like(pts)
        self._check(getattr(self._lib, "wtp_normals_" + _sfx(pts.dtype))(self._h, _vp(pts), C.c_int64(pts.shape[0]), C.c_int32(pts.shape[1]),
                                                                       C.c_int32(k), _vp(out)))
        return out

    def gradient_limit(self, centers, h0, g: float, k: int = 12, tol: float = 1.0e-3, max_sweeps: int = 2000):
        """_gradient_limit_field (src/discretization/algorithms/octree.jl:677-717) on the leaf centres -> (h, sweeps)."""
        centers = _as_points(centers)
        h0 = np.ascontiguousarray(h0, dtype=centers.dtype)
        out = np.empty_like(h0)
        sweeps = C.c_int32(0)
        sfx = _sfx(centers.dtype)
        gg = C.c_float(g) if sfx == "f32" else C.c_double(g)
        self._check(getattr(self._lib, "wtp_gradient_limit_" + sfx)(self._h, _vp(centers), C.c_int64(centers.shape[0]), C.c_int32(centers.shape[1]),
                                                                   _vp(h0), gg, C.c_int32(k), C.c_double(tol), C.c_int32(max_sweeps), _vp(out), C.byref(sweeps)))
        return out, int(sweeps.value)

    def metrics(self, pts, k=20) -> dict:
        pts = _as_points(pts)
        out = CloudMetrics()
        self._check(getattr(self._lib, "wtp_metrics_" + _sfx(pts.dtype))(
            self._h, _vp(pts), C.c_int64(pts.shape[0]), C.c_int32(pts.shape[1]), C.c_int32(k), C.byref(out)))
        return {n: getattr(out, n) for n, _ in CloudMetrics._fields_}


_default_ctx: Context | None = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("WTP_USE_LOCAL_RANK") else 0)
    return _default_ctx
