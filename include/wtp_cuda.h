/*
 * wtp_cuda.h — C ABI of libwtp_cuda.so, the B200 (sm_100a) replacement for the
 * data-parallel hot path of WhatsThePoint.jl.
 *
 * The reference has no FFI; the boundary is created by overriding the three
 * internal Julia functions that are the only call sites of the third-party
 * KD-tree on this path (paths relative to the reference checkout):
 *
 *   _build_knn_neighbors(points, k)          src/topology.jl:79-84    -> wtp_knn_{f32,f64}
 *   _build_radius_neighbors(points, radius)  src/topology.jl:91-97    -> wtp_radius_count/fill_{f32,f64}
 *   _relax!(p, p_old, snap, spacing, ...)    src/repel.jl:202-339     -> wtp_repel_{f32,f64}
 *
 * plus the sibling consumers of the same k-NN primitive:
 *
 *   searchdists(cloud, KNearestSearch)       src/neighbors.jl:16-21   -> wtp_knn_self_{f32,f64}
 *   spacing.(points)                         src/repel.jl:61,209,251  -> wtp_spacing_eval_{f32,f64}
 *   metrics / spacing_metrics / ..._fidelity src/metrics.jl:19-129    -> wtp_metrics_{f32,f64}
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ types, no exceptions cross the ABI.
 *   - every entry point returns a wtp_status (0 = OK); the message for the last
 *     failure on a context is wtp_last_error(ctx).
 *   - point sets are N x D row-major arrays of T (D = 2 or 3): exactly the bytes of
 *     a Julia Vector{Point{𝔼{D},Cartesian{...,Quantity{T}}}} (units are type-level).
 *   - indices crossing the ABI are int64, 1-based, in the caller's point order.
 *   - neighbour order is canonical: ascending (d2, index), d2 = ((dx*dx + dy*dy) + dz*dz)
 *     evaluated in T without fused multiply-add.
 *   - host entry points take HOST pointers and do their own H2D/D2H; the *_dev entry
 *     points take DEVICE pointers (same layouts) and enqueue on the context stream.
 *   - the caller owns every buffer it passes; the library owns all device memory in
 *     the context and keeps no host pointer after a call returns.
 *   - one context is not re-entrant. There is no CPU fallback: without a usable
 *     CUDA device wtp_create fails with WTP_ERR_CUDA.
 */
#ifndef WTP_CUDA_H
#define WTP_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wtp_ctx wtp_ctx;

typedef enum {
    WTP_OK = 0,
    WTP_ERR_BAD_ARG = 1,       /* null pointer, D not in {2,3}, negative size ...       -> ArgumentError */
    WTP_ERR_K_TOO_LARGE = 2,   /* k+1 > N (upstream knn throws), or k above WTP_MAX_K   -> ArgumentError */
    WTP_ERR_UNSUPPORTED = 3,   /* user force model / spacing callable / constrain closure -> ErrorException */
    WTP_ERR_CUDA = 4,          /* CUDA runtime failure (message has the CUDA error)     */
    WTP_ERR_NCCL = 5,
    WTP_ERR_OOM = 6,
    WTP_ERR_STATE = 7          /* e.g. wtp_radius_fill without a preceding count        */
} wtp_status;

/* Largest list length the warp-resident top-k holds: k+1 for topology and search (up to 8 list rows per lane), k for
 * repel (up to 4). Beyond that the call fails with WTP_ERR_K_TOO_LARGE (no fallback). */
#define WTP_MAX_K 256
#define WTP_MAX_K_REPEL 128

/* ------------------------------------------------------------------ context */

/* device: CUDA ordinal. Fails (no fallback) if the device is unusable. */
int32_t wtp_create(wtp_ctx** out, int32_t device);
/* One context over n_devices GPUs of the box, in ONE process (for a host program that is a single process, like a Julia
 * session): the library runs one host thread and one stream per device, the devices are the ranks of one NCCL
 * communicator with peer access between every pair. The HOST entry points that shard — wtp_knn_*, wtp_knn_self_*,
 * wtp_radius_count_* / wtp_radius_fill, wtp_repel_* with the identity wall — are answered by all devices together and
 * fill the caller's arrays completely, exactly as the same call on a single-device context would (k-NN: every device
 * answers a run of the sorted order and the rows travel over NVLink to the device that owns their caller range; radius:
 * contiguous caller ranges, one CSR; repel: runs of the sorted order, every device ends with the whole state). Everything
 * else (metrics, isinside, mesh queries, normals, ..., the mesh-wall repel, point sets of fewer than 4096 per device)
 * runs on the first device. Device-pointer entry points and wtp_set_stream are refused (WTP_ERR_UNSUPPORTED).
 * wtp_comm_rank / wtp_comm_world of the handle are 0 / 1: it is one logical context. n_devices == 1 is wtp_create. */
int32_t wtp_create_multi(wtp_ctx** out, const int32_t* devices, int32_t n_devices);
void wtp_destroy(wtp_ctx* ctx);
const char* wtp_last_error(const wtp_ctx* ctx);
const char* wtp_status_string(int32_t status);
int32_t wtp_version(void);

/* Use an existing CUDA stream (cudaStream_t passed as void*) for all work of this
 * context; NULL restores the context's own stream. */
int32_t wtp_set_stream(wtp_ctx* ctx, void* cuda_stream);

/* Pin / unpin a host range so H2D/D2H of the host entry points run at full PCIe rate. */
int32_t wtp_host_register(wtp_ctx* ctx, void* ptr, int64_t bytes);
int32_t wtp_host_unregister(wtp_ctx* ctx, void* ptr);

/* Tuning knob: target number of points per grid cell (default 0.36*K in 3-D, 0.45*K in 2-D for a
 * list of K entries, i.e. 8 / 10 at k = 21; <= 0 restores the default). */
int32_t wtp_set_cell_occupancy(wtp_ctx* ctx, double points_per_cell);

/* Multi-GPU, one process per GPU. The NCCL unique id (128 bytes) is created on one
 * rank with wtp_comm_unique_id and distributed by the host program; a null id makes a shard-only
 * context (rank and world set, no communicator): enough for k-NN and radius, refused by repel
 * (WTP_ERR_STATE). After wtp_comm_init:
 *  - k-NN: every rank answers the contiguous run [wtp_shard_begin(N), wtp_shard_end(N)) of the
 *    SPATIALLY SORTED order, no collective. The rank indexes only the layers of the grid that run
 *    needs (its own plus two on either side); if a search ever has to leave them (strongly graded
 *    clouds) the call is repeated on the whole index and the context stops windowing.
 *    Device entry points write a compact wtp_shard_owned_count() x k table; wtp_shard_owned gives
 *    the caller index of each row. Host entry points fill rows of the caller's N x k table (the
 *    others stay untouched) and wtp_shard_owned says which: with a communicator (non-null id) whose
 *    ranks can map each other's memory, the kernels hand every row to the rank that owns its CALLER
 *    range over NVLink, and each rank fills the contiguous rows [wtp_shard_begin(N), wtp_shard_end(N))
 *    of the table (written as int64 by the DMA engine itself when the table is pinned); otherwise
 *    (shard-only context, distances requested) the rows of the rank's sorted run, wherever they are.
 *  - radius: the contiguous range [wtp_shard_begin, wtp_shard_end) of the caller's order.
 *  - wtp_repel_*: the same kind of range of the movable points, all-gathering the moved
 *    positions over NCCL every iteration. */
int32_t wtp_comm_unique_id(void* out128);
int32_t wtp_comm_init(wtp_ctx* ctx, int32_t rank, int32_t world, const void* unique_id128);
int32_t wtp_comm_rank(const wtp_ctx* ctx);
int32_t wtp_comm_world(const wtp_ctx* ctx);
/* Contiguous block partition of n items: rank r owns [begin, end). */
int64_t wtp_shard_begin(int64_t n, int32_t rank, int32_t world);
int64_t wtp_shard_end(int64_t n, int32_t rank, int32_t world);
/* Rows answered by the last k-NN call of a sharded context: their number, and their caller
 * indices (int64, 1-based) in the order of the compact device table. */
int64_t wtp_shard_owned_count(const wtp_ctx* ctx);
int32_t wtp_shard_owned(wtp_ctx* ctx, int64_t* ids /* host */);
int32_t wtp_shard_owned_dev(wtp_ctx* ctx, int64_t* d_ids);

/* ------------------------------------------------------------ instrumentation */

typedef struct {
    float ms_h2d;        /* host entry points only */
    float ms_bbox;       /* bounding box + grid parameters */
    float ms_cellkey;    /* cell-key kernel */
    float ms_sort;       /* radix sort (all passes) */
    float ms_reorder;    /* gather into sorted float4 tiles + cell starts */
    float ms_query;      /* k-NN query / radius count+fill / repel sweep */
    float ms_scan;       /* exclusive scan (radius CSR); spacing evaluation (repel) */
    float ms_reduce;     /* repel reductions + stop-test scalars */
    float ms_comm;       /* NCCL all-gather (multi-GPU repel) */
    float ms_d2h;        /* host entry points only */
    float ms_total;
    int32_t sort_passes; /* radix passes actually run (8-bit digits) */
    int32_t query_launches;
    int64_t n_cells;
    int64_t n_ring_expanded; /* queries that needed more than the 3^D block */
    /* queries the tiled front end handed to the general kernel, by reason: block too sparse / too dense for
     * the tile / look-alike keys or exact ties (k-NN: last call; repel: last iteration) */
    int64_t n_leftover_sparse, n_leftover_dense, n_leftover_other;
    /* sharded k-NN: points in the window of the grid this rank indexed (0: the whole set was indexed), and the
     * queries whose search left the window (> 0: the call was repeated on the whole index) */
    int64_t n_window_points, n_window_missed;
    /* sharded repel: ranks whose run buffers the sweep kernels wrote directly over NVLink peer memory (0: the runs
     * were exchanged with an NCCL all-gather after the sweep); sharded host k-NN: ranks taking part in the row exchange */
    int64_t n_peer_ranks;
    /* host k-NN entry points: bytes the call moved over PCIe (points in; rows out: 4 bytes per entry when they are widened
     * on the host, 8 when the DMA engine writes the int64 table itself, plus the caller indices of a sharded scatter) */
    int64_t bytes_h2d, bytes_d2h;
} wtp_timing;

/* enable != 0: record CUDA events around each phase of subsequent calls. */
int32_t wtp_set_timing(wtp_ctx* ctx, int32_t enable);
/* Timing of the most recent call (synchronises the context stream). */
int32_t wtp_get_timing(wtp_ctx* ctx, wtp_timing* out);
/* Number of kernels this context has launched since creation. */
int64_t wtp_launch_count(const wtp_ctx* ctx);

/* ---------------------------------------------------------------- topology */

/* _build_knn_neighbors (src/topology.jl:79-84): for every point the k nearest OTHER
 * points. Queries k+1, orders by (d2, index), drops position 1 (the reference's
 * n[2:end]). out_idx: N x k int64, 1-based. out_dist (nullable): N x k distances
 * sqrt(d2) in T. Requires N >= k+1. Sharded contexts: see wtp_comm_init above. */
int32_t wtp_knn_f32(wtp_ctx*, const float* pts, int64_t N, int32_t D, int32_t k,
                    int64_t* out_idx, float* out_dist);
int32_t wtp_knn_f64(wtp_ctx*, const double* pts, int64_t N, int32_t D, int32_t k,
                    int64_t* out_idx, double* out_dist);
int32_t wtp_knn_dev_f32(wtp_ctx*, const float* d_pts, int64_t N, int32_t D, int32_t k,
                        int64_t* d_out_idx, float* d_out_dist);
int32_t wtp_knn_dev_f64(wtp_ctx*, const double* d_pts, int64_t N, int32_t D, int32_t k,
                        int64_t* d_out_idx, double* d_out_dist);
/* The same table as 4-byte indices (N < 2^31): for a caller that keeps the adjacency on the device or backs its row type
 * with Int32 storage (an AbstractVector{Int} row converting on getindex satisfies KNNTopology, src/topology.jl:25). The
 * kernels write half the bytes; contents and order are those of the int64 table. Device pointers only. */
int32_t wtp_knn_dev_i32_f32(wtp_ctx*, const float* d_pts, int64_t N, int32_t D, int32_t k,
                            int32_t* d_out_idx, float* d_out_dist);
int32_t wtp_knn_dev_i32_f64(wtp_ctx*, const double* d_pts, int64_t N, int32_t D, int32_t k,
                            int32_t* d_out_idx, double* d_out_dist);

/* search / searchdists (src/neighbors.jl:9-21): the k nearest points INCLUDING the
 * query's own entry (position 1 whenever the point is not duplicated). */
int32_t wtp_knn_self_f32(wtp_ctx*, const float* pts, int64_t N, int32_t D, int32_t k,
                         int64_t* out_idx, float* out_dist);
int32_t wtp_knn_self_f64(wtp_ctx*, const double* pts, int64_t N, int32_t D, int32_t k,
                         int64_t* out_idx, double* out_dist);

/* _build_radius_neighbors (src/topology.jl:91-97): two calls so the caller allocates
 * the exact CSR. count: offsets[N+1] (0-based exclusive prefix, offsets[N] = nnz).
 * fill: indices[nnz], 1-based, each row ascending by index, self removed BY INDEX,
 * inclusion test d2 <= r*r in T. fill must directly follow count on the same ctx. */
int32_t wtp_radius_count_f32(wtp_ctx*, const float* pts, int64_t N, int32_t D, float r,
                             int64_t* offsets);
int32_t wtp_radius_count_f64(wtp_ctx*, const double* pts, int64_t N, int32_t D, double r,
                             int64_t* offsets);
int32_t wtp_radius_fill(wtp_ctx*, int64_t* indices);
int32_t wtp_radius_count_dev_f32(wtp_ctx*, const float* d_pts, int64_t N, int32_t D, float r,
                                 int64_t* d_offsets);
int32_t wtp_radius_count_dev_f64(wtp_ctx*, const double* d_pts, int64_t N, int32_t D, double r,
                                 int64_t* d_offsets);
int32_t wtp_radius_fill_dev(wtp_ctx*, int64_t* d_indices);
/* nnz of the pending count (host copy of offsets[N]); -1 if none. */
int64_t wtp_radius_nnz(const wtp_ctx*);

/* -------------------------------------------------------------------- repel */

/* Random directions. _safe_direction (src/repel.jl:358-364) gives a coincident neighbour (r == 0) a random unit
 * vector, randn(...)/norm, drawn from Julia's global RNG; no other implementation can reproduce that stream, so the
 * library defines its own, counter-based and stateless: with
 *     mix64(z): z += 0x9e3779b97f4a7c15; z = (z ^ z>>30) * 0xbf58476d1ce4e5b9; z = (z ^ z>>27) * 0x94d049bb133111eb; z ^ z>>31
 *     key = mix64(kick_seed ^ mix64(iteration)),  h1 = mix64(key ^ (i << 32 | j))        (i, j: snapshot-global, 0-based)
 * the direction of neighbour j seen from point i is the first v_c, c = 0, 1, ..., with 2^-10 <= |v_c|^2 <= 1, divided
 * by |v_c|, where v_c[d] = (int(mix64(h1 + c) >> 21 d & 0x1fffff) - 2^20) * 2^-20 — uniform on the circle / sphere.
 * Only exactly rounded operations: the oracle (oracle/wtp_oracle.cpp) produces the same bits. */

/* Force laws F(u), u = r/s (src/repel_forces.jl). */
enum { WTP_FORCE_INVERSE = 0,      /* 1/(u^2+beta)^2                      :37      */
       WTP_FORCE_EQUILIBRIUM = 1,  /* (1-u^2)/(u^2+beta)^2                :57-60   */
       WTP_FORCE_CLIPPED = 2,      /* max((u0^2-u^2)/(u^2+beta)^2, 0)     :96-100  */
       WTP_FORCE_STRONG = 3 };     /* (1-u^2)/(u^2+beta)^gamma            :124-127 */
typedef struct {
    int32_t kind;
    double beta, u0, gamma;        /* converted to T at the boundary */
} wtp_force;

/* Spacing callables (src/discretization/spacings.jl). */
enum { WTP_SPACING_CONSTANT = 0,        /* a = dx                                  :35-39   */
       WTP_SPACING_LOGLIKE = 1,         /* a = base_size, b = growth_rate          :67-72   */
       WTP_SPACING_BOUNDARY_LAYER = 2 };/* a = at_wall, b = bulk, c = thickness    :121-133 */
typedef struct {
    int32_t kind;
    double a, b, c;
    const void* bnd_pts;           /* the spacing's own boundary point set (T, D), HOST pointer */
    int64_t n_bnd;                 /* (DEVICE pointer for the *_dev entry points)               */
} wtp_spacing;

enum { WTP_WALL_IDENTITY = 0,      /* repel(cloud, spacing): src/repel.jl:82         */
       WTP_WALL_MESH = 1 };        /* repel(cloud, spacing, octree): :158-160,448-469 */

/* Triangle mesh for the wall rule of the 3-argument repel (3-D only): the arrays of the
 * reference's TriangleIndex (src/octree/triangle_octree.jl:22-31) flattened per triangle. The
 * library builds its own device search structure (Morton BVH + a cell-class grid); results equal
 * the reference's octree queries: nearest triangle by (d2, triangle index), closest point by
 * Ericson's region tests (src/octree/geometric_utils.jl:68-136), inside test = sign of
 * dot(p - cp, pseudonormal of the closest feature) < 0 inside the mesh bounding box
 * (src/octree/triangle_octree.jl:71-99, 583-607). HOST pointers, coordinates in the cloud's T. */
typedef struct {
    const void* triangles;         /* n_tri x 9 of T: v1, v2, v3                                  */
    const void* feature_normals;   /* n_tri x 21 of T: face, vertex 1..3, edge 12, 13, 23          */
                                   /*   (index.face / index.vertex / index.edge pseudonormals)    */
    int64_t n_tri;
    double bbox_min[3], bbox_max[3]; /* index.bbox_min / bbox_max (domain_bounds)                 */
    double offset_dist;            /* 1e-6 * |bbox_max - bbox_min|  (src/repel.jl:150)            */
    const uint8_t* is_bnd;         /* n_move flags (src/repel.jl:155); rewritten when deposit_ratio > 0 */
    int64_t* tri_indices;          /* out, n_move, 1-based landing triangle (0 = untouched)       */
    uint8_t* escaped;              /* out, n_move                                                 */
} wtp_wall_mesh;

typedef struct {
    int32_t k;                     /* neighbourhood size INCLUDING self (default 21)      */
    int32_t max_iters;
    int32_t rebuild_every;         /* >= 1 (src/repel.jl:74)                              */
    int32_t stall_after;
    int32_t kick_after;            /* > 0: _maybe_kick! (src/repel.jl:415-433) with the library's own
                                      random stream (kick_seed): the reference draws from Julia's global
                                      RNG, so trajectories after a kick are not reproducible across the two */
    int32_t wall;                  /* WTP_WALL_*                                          */
    int32_t want_trace;
    int32_t reserved;
    double alpha_lo, alpha_max;    /* ustrip(α_min), ustrip(α) (src/repel.jl:86)          */
    double tol, cv_target;
    int64_t n_protected;           /* snapshot-global count of points a kick avoids: n_boundary (:85, :172) */
    uint64_t kick_seed;            /* seed of the library's random stream: the kick directions and the unit vectors of
                                      coincident pairs ("random directions" below)                                */
    double deposit_ratio;          /* > 0 (mesh wall, n_fixed = 0, one GPU): _deposit_escaped! after every sweep
                                      (src/repel.jl:161-168, 328, 483-514): escaped volume points are projected onto
                                      their nearest triangle and become boundary points unless a boundary point
                                      already sits within deposit_ratio * spacing of the landing site. Projection,
                                      spacing and the snapshot k-NN of the landing sites run on the device, batched;
                                      the occupancy sweep is serial by design and runs on the host in index order.
                                      wall->is_bnd is then an in/out array. */
} wtp_repel_params;

enum { WTP_STOP_MAX_ITERS = 0, WTP_STOP_TOL = 1, WTP_STOP_CV_TARGET = 2, WTP_STOP_STALL = 3 };
typedef struct {
    int32_t iters;                 /* length of conv                                      */
    int32_t stop_reason;           /* WTP_STOP_*                                          */
    double last_cv;                /* d_NN/s CV of the last monitored sweep (NaN if off)  */
} wtp_repel_result;

/* One closest-pair record per iteration (src/repel.jl:294-296, 396-403). */
typedef struct {
    double r, s, r_over_s;
    int64_t idx_a, idx_b;          /* snapshot-global, 1-based, idx_a < idx_b */
} wtp_trace_entry;

/* _relax! (src/repel.jl:202-339). snap = (n_fixed + n_move) x D of T: static head,
 * movable tail, the tail is updated in place with the final positions (pre-sweep
 * positions on a cv_target stop). conv has max_iters slots, trace (nullable) too. */
int32_t wtp_repel_f32(wtp_ctx*, float* snap, int64_t n_fixed, int64_t n_move, int32_t D,
                      const wtp_spacing*, const wtp_force*, const wtp_repel_params*,
                      const wtp_wall_mesh* wall /*nullable*/, float* conv,
                      wtp_trace_entry* trace /*nullable*/, wtp_repel_result* result);
int32_t wtp_repel_f64(wtp_ctx*, double* snap, int64_t n_fixed, int64_t n_move, int32_t D,
                      const wtp_spacing*, const wtp_force*, const wtp_repel_params*,
                      const wtp_wall_mesh* wall /*nullable*/, double* conv,
                      wtp_trace_entry* trace /*nullable*/, wtp_repel_result* result);
/* Device-resident variant: d_snap and spacing->bnd_pts are DEVICE pointers; conv, trace,
 * result stay HOST pointers (a few scalars per iteration). Wall must be identity. */
int32_t wtp_repel_dev_f32(wtp_ctx*, float* d_snap, int64_t n_fixed, int64_t n_move, int32_t D,
                          const wtp_spacing*, const wtp_force*, const wtp_repel_params*,
                          float* conv, wtp_trace_entry* trace, wtp_repel_result* result);
int32_t wtp_repel_dev_f64(wtp_ctx*, double* d_snap, int64_t n_fixed, int64_t n_move, int32_t D,
                          const wtp_spacing*, const wtp_force*, const wtp_repel_params*,
                          double* conv, wtp_trace_entry* trace, wtp_repel_result* result);

/* spacing.(points) for the three built-in spacings (used for the default
 * α = minimum(spacing.(to(cloud)))/20, src/repel.jl:61, and by the metrics). */
int32_t wtp_spacing_eval_f32(wtp_ctx*, const wtp_spacing*, const float* pts, int64_t N,
                             int32_t D, float* out);
int32_t wtp_spacing_eval_f64(wtp_ctx*, const wtp_spacing*, const double* pts, int64_t N,
                             int32_t D, double* out);

/* Batched geometry queries on the same mesh structure (is_bnd / tri_indices / escaped unused):
 * isinside(points, octree) (src/octree/triangle_octree.jl:97-115) -> out[i] in {0,1};
 * _project_to_boundary (src/repel.jl:522-537) -> out_pts N x 3, out_tri 1-based (0: none). */
int32_t wtp_mesh_isinside_f32(wtp_ctx*, const wtp_wall_mesh*, const float* pts, int64_t N, uint8_t* out);
int32_t wtp_mesh_isinside_f64(wtp_ctx*, const wtp_wall_mesh*, const double* pts, int64_t N, uint8_t* out);
int32_t wtp_mesh_project_f32(wtp_ctx*, const wtp_wall_mesh*, const float* pts, int64_t N, float* out_pts, int64_t* out_tri);
int32_t wtp_mesh_project_f64(wtp_ctx*, const wtp_wall_mesh*, const double* pts, int64_t N, double* out_pts, int64_t* out_tri);

/* isinside(points, cloud) against the boundary POINT cloud (src/isinside.jl), the survivor filter of
 * repel(cloud, spacing) (src/repel.jl:90). D = 3: Green's-function sum over the boundary elements
 * (positions M x 3, unit normals M x 3, areas M; all surfaces concatenated), inside iff the sum < -2*pi
 * (:86-106). D = 2: winding number over the M polygon points, which must be ordered around the boundary
 * (:18-35; WTP_ERR_BAD_ARG when M < 3 or the signed area vanishes, :37-69); normals and areas are unused.
 * out: N flags; sums (nullable): the raw N sums in T. HOST pointers. */
int32_t wtp_isinside_f32(wtp_ctx*, const float* pts, int64_t N, int32_t D, const float* bnd_pts, const float* bnd_normals,
                         const float* bnd_areas, int64_t M, uint8_t* out, float* sums);
int32_t wtp_isinside_f64(wtp_ctx*, const double* pts, int64_t N, int32_t D, const double* bnd_pts, const double* bnd_normals,
                         const double* bnd_areas, int64_t M, uint8_t* out, double* sums);

/* compute_force(model, u) elementwise on the device (src/repel_forces.jl:22). */
int32_t wtp_force_eval_f32(wtp_ctx*, const wtp_force*, const float* u, int64_t n, float* out);
int32_t wtp_force_eval_f64(wtp_ctx*, const wtp_force*, const double* u, int64_t n, double* out);

/* ------------------------------------------------------------------ metrics */

/* metrics(cloud; k) (src/metrics.jl:19-41): statistics of the distances to the k-1
 * nearest other points (the reference queries k including self and drops it). */
typedef struct {
    double avg, std, max, min;         /* means over points of per-point mean/std/max/min */
    double separation, fill, mesh_ratio;
} wtp_cloud_metrics;
int32_t wtp_metrics_f32(wtp_ctx*, const float* pts, int64_t N, int32_t D, int32_t k,
                        wtp_cloud_metrics* out);
int32_t wtp_metrics_f64(wtp_ctx*, const double* pts, int64_t N, int32_t D, int32_t k,
                        wtp_cloud_metrics* out);

/* _near_duplicate_keep_mask(pts, spacings, ratio) (src/repel.jl:565-580), the cull of repel(...; cull_ratio):
 * greedy, order-preserving: a point is dropped when a kept, lower-indexed... more precisely, walking the points in
 * index order, every kept point i drops each still-kept j != i with |p_j - p_i| < ratio * spacings[i]. The ball search
 * at radius ratio * max(spacings) runs on the device (radius CSR); the index-ordered sweep is serial by design and runs
 * on the host over that CSR. keep: N flags (1 = kept). HOST pointers. */
int32_t wtp_cull_mask_f32(wtp_ctx*, const float* pts, int64_t N, int32_t D, const float* spacings, double ratio, uint8_t* keep);
int32_t wtp_cull_mask_f64(wtp_ctx*, const double* pts, int64_t N, int32_t D, const double* spacings, double ratio, uint8_t* keep);

/* spacing_metrics(cloud, spacing; k) (src/metrics.jl:56-71): error_i = |mean distance to the k-1 nearest
 * others - s(x_i)| / s(x_i); max, mean and (sample) standard deviation over the points. */
typedef struct { double max_error, mean_error, std_error; } wtp_spacing_metrics_t;
int32_t wtp_spacing_metrics_f32(wtp_ctx*, const float* pts, int64_t N, int32_t D, int32_t k, const wtp_spacing*, wtp_spacing_metrics_t* out);
int32_t wtp_spacing_metrics_f64(wtp_ctx*, const double* pts, int64_t N, int32_t D, int32_t k, const wtp_spacing*, wtp_spacing_metrics_t* out);

/* spacing_fidelity_metrics(cloud, spacing; k, coord_radius) (src/metrics.jl:88-129): u_i = d_NN(i) / h(x_i)
 * (nearest OTHER point by index among the k nearest); mean, cv = std/mean, the 5/50/95 % quantiles (linear
 * interpolation between order statistics, Julia's default), and the mean number of neighbours within
 * coord_radius * h(x_i). k is clamped to N. */
typedef struct { double mean_dnn_h, cv, p05, p50, p95, coordination; } wtp_spacing_fidelity_t;
int32_t wtp_spacing_fidelity_f32(wtp_ctx*, const float* pts, int64_t N, int32_t D, int32_t k, double coord_radius, const wtp_spacing*,
                                 wtp_spacing_fidelity_t* out);
int32_t wtp_spacing_fidelity_f64(wtp_ctx*, const double* pts, int64_t N, int32_t D, int32_t k, double coord_radius, const wtp_spacing*,
                                 wtp_spacing_fidelity_t* out);

/* ------------------------------------------- other consumers of the k-NN index */

/* compute_normals(points; k) (src/normals.jl:9-44, _compute_normal :65-70): per point the unit eigenvector of the
 * smallest eigenvalue of cov(its k nearest points, itself included) — the PCA normal of Hoppe (1992). k is clamped to
 * N (:16). The sign is not defined by the reference (LAPACK's; orient_normals! fixes it afterwards): here the first
 * nonzero component is positive. out: N x D of T. HOST pointers. */
int32_t wtp_normals_f32(wtp_ctx*, const float* pts, int64_t N, int32_t D, int32_t k, float* out);
int32_t wtp_normals_f64(wtp_ctx*, const double* pts, int64_t N, int32_t D, int32_t k, double* out);

/* _gradient_limit_field(node_tree, leaves, h0_field, g; k, tol, max_sweeps) (src/discretization/algorithms/octree.jl:
 * 677-717) on the leaf centres: the g-Lipschitz envelope of h0 by Jacobi min-plus sweeps over the k-NN graph of the
 * centres, h[a] <- min(h[a], min_j h[j] + g * d_aj), until the largest relative change of a sweep is below tol (or
 * max_sweeps). centers: n x D, h0 / out: n values (the caller maps them to and from its box-indexed field), sweeps
 * (nullable): the number of sweeps run. k is clamped to n. HOST pointers. */
int32_t wtp_gradient_limit_f32(wtp_ctx*, const float* centers, int64_t n, int32_t D, const float* h0, float g, int32_t k,
                               double tol, int32_t max_sweeps, float* out, int32_t* sweeps);
int32_t wtp_gradient_limit_f64(wtp_ctx*, const double* centers, int64_t n, int32_t D, const double* h0, double g, int32_t k,
                               double tol, int32_t max_sweeps, double* out, int32_t* sweeps);

#ifdef __cplusplus
}
#endif
#endif /* WTP_CUDA_H */
