// repel.cu — the node-repulsion relaxation: _relax! (src/repel.jl:202-339).
//
// One fused sweep kernel per iteration (src/repel.jl:256-292): warp-per-point k-NN on the
// snapshot grid, force accumulation in ascending (d2, index) order, adaptive step with
// the one-spacing cap, wall rule, position update, and the block-reduced stopping
// metrics (max |F|*s  :293, d_NN/s moments  :374-386, closest pair  :396-403) plus the
// bounding box of the moved points for the next grid. A tiny second kernel folds the
// per-CTA partials in a fixed order, the host replays the stop logic (:305-337).
#include <cmath>
#include <cstdlib>
#include <limits>
#include <random>
#include <algorithm>
#include <vector>

#include "kernels.cuh"
#include "knn_core.cuh"
#include "knn_tile.cuh"

namespace wtp {

#define LAUNCH_CHECK(ctx)                         \
    do {                                          \
        (ctx)->launches++;                        \
        WTP_CUDA_CHECK(cudaPeekAtLastError());    \
    } while (0)

// comm.cu
bool comm_peer_buffers(wtp_ctx* ctx, PeerSet& pb, size_t bytes_each);
void comm_allgather_rows(wtp_ctx* ctx, void* d_buf, int64_t n_rows, size_t row_bytes);
void comm_allgather_fixed(wtp_ctx* ctx, const void* d_in, void* d_out, size_t bytes_per_rank);

// ------------------------------------------------------------------ forces
// src/repel_forces.jl:37, 57-60, 96-100, 124-127. Every law depends on u = r/s through u^2 only, so the sweeps
// evaluate F(u^2) with u^2 = d2 / s^2 and never need u itself; with the direction (xi - xj)/r the term of one
// neighbour is F(u^2) * rsqrt(d2) * (xi - xj): one square root and one division per neighbour. This translation
// unit is compiled with FMA contraction: the north star's contract for repel is a tolerance (1e-6 s in Float64,
// 1e-3 s in Float32 after 10 iterations), not bit equality; what stays canonical is the neighbour SELECTION
// (keys from dist2_rn, explicit round-to-nearest operations, ordered by (d2, index)).
// FK: the law known at compile time (the default ClippedSpacingForce), or -1 = a switch on f.kind.
template <class T> __device__ __forceinline__ T fast_div(T a, T b);
template <> __device__ __forceinline__ float fast_div<float>(float a, float b) { return __fdividef(a, b); }   // MUFU.RCP + FMUL, 2 ulp
template <> __device__ __forceinline__ double fast_div<double>(double a, double b) { return a / b; }
template <class T> __device__ __forceinline__ T fast_sqrt(T a);
template <> __device__ __forceinline__ float fast_sqrt<float>(float a) { return __fsqrt_rn(a); }
template <> __device__ __forceinline__ double fast_sqrt<double>(double a) { return sqrt(a); }

template <class T>
__device__ __noinline__ T force_u2_generic(const ForceP<T> f, T u2) {
    switch (f.kind) {
        case WTP_FORCE_INVERSE: { const T t = u2 + f.beta; return fast_div<T>((T)1, t * t); }
        case WTP_FORCE_EQUILIBRIUM: { const T t = u2 + f.beta; return fast_div<T>((T)1 - u2, t * t); }
        case WTP_FORCE_CLIPPED: { const T t = u2 + f.beta; const T F = fast_div<T>(f.u0 * f.u0 - u2, t * t); return F > (T)0 ? F : (T)0; }
        default: return ((T)1 - u2) / pow(u2 + f.beta, f.gamma);
    }
}
// the term of one neighbour at squared distance d2 > 0, as the weight w of (xi - xj): F(d2 / s^2) / sqrt(d2).
// Returns 0 when the law gives no force there (the clipped law beyond u0).
template <class T, int FK>
__device__ __forceinline__ T pair_weight(const ForceP<T>& f, T u0sq, T d2, T inv_s2) {
    const T u2 = d2 * inv_s2;
    if (FK == WTP_FORCE_CLIPPED) {
        const T q = u0sq - u2;
        if (!(q > (T)0)) return (T)0;
        const T t = u2 + f.beta;
        return fast_div<T>(q, t * t * fast_sqrt<T>(d2));
    }
    return fast_div<T>(force_u2_generic<T>(f, u2), fast_sqrt<T>(d2));
}
// F(0): the magnitude of the push between coincident points (_safe_direction, src/repel.jl:358-364)
template <class T, int FK>
__device__ __forceinline__ T force_at_zero(const ForceP<T>& f, T u0sq) {
    if (FK == WTP_FORCE_CLIPPED) { const T F = fast_div<T>(u0sq, f.beta * f.beta); return F > (T)0 ? F : (T)0; }
    return force_u2_generic<T>(f, (T)0);
}

// ------------------------------------------------------------------- sweep
template <class T> inline T t_inf_host() { return std::numeric_limits<T>::infinity(); }
template <class T> __host__ __device__ inline T t_max();
template <> __host__ __device__ inline float t_max<float>() { return 3.402823466e+38f; }
template <> __host__ __device__ inline double t_max<double>() { return 1.7976931348623157e+308; }

template <class T>
struct RepelPartial {
    double s1, s2;              // sum u, sum u*u, u = d_NN / s     (_dnn_cv, :374-386)
    unsigned long long n;       // points swept
    T max_force;                // max |F|*s                        (:293)
    T min_nn;                   // closest pair                     (:396-403)
    uint32_t min_id;            // movable id of the pair's first point
    uint32_t min_nn_idx;        // snapshot-global 0-based index of its nearest neighbour
    T lo[3], hi[3];             // bounding box of the new positions
    unsigned long long missed;  // searches that had to leave the window of a windowed index (the iteration is redone)
};

template <class T>
__device__ __forceinline__ void partial_init(RepelPartial<T>& p) {
    p.s1 = 0; p.s2 = 0; p.n = 0; p.missed = 0; p.max_force = (T)0; p.min_nn = t_inf<T>(); p.min_id = 0xffffffffu; p.min_nn_idx = 0xffffffffu;
    for (int d = 0; d < 3; ++d) { p.lo[d] = t_inf<T>(); p.hi[d] = -t_inf<T>(); }
}
// fold b into a; b covers later points than a (sum order = point order within a CTA)
template <class T>
__host__ __device__ inline void partial_merge(RepelPartial<T>& a, const RepelPartial<T>& b) {
    a.s1 += b.s1; a.s2 += b.s2; a.n += b.n; a.missed += b.missed;
    a.max_force = b.max_force > a.max_force ? b.max_force : a.max_force;
    if (b.min_nn < a.min_nn || (b.min_nn == a.min_nn && b.min_id < a.min_id)) { a.min_nn = b.min_nn; a.min_id = b.min_id; a.min_nn_idx = b.min_nn_idx; }
    for (int d = 0; d < 3; ++d) { a.lo[d] = b.lo[d] < a.lo[d] ? b.lo[d] : a.lo[d]; a.hi[d] = b.hi[d] > a.hi[d] ? b.hi[d] : a.hi[d]; }
}

template <class T>
struct SweepArgs {
    Grid<T> g;
    const P4<T>* sorted;
    const uint32_t* cell_start;
    const T* S;            // snapshot coordinates, n_all x D, caller order
    const T* P_old;        // n_move x D
    T* P_new;              // n_move x D
    const T* s_cur;        // spacing at P_old per movable point (null: constant)
    const T* spacings;     // n_all, snapshot-global, as of the last rebuild
    T s_const;
    uint32_t n_fixed, n_all, id_lo, id_hi;
    const uint32_t* qlist; // sorted positions to sweep (null: all, filtered by id range)
    uint32_t nq;
    const uint32_t* nq_dev; // device-side length of qlist (the tiled sweep's leftovers), or null
    int kk, rebuild;
    // Sharded by runs of the sorted order: this rank sweeps the sorted positions [s_begin, s_end) and writes the new
    // position of position j, with the point's caller index, to Cp[r][j - s_begin] for r < n_peers: its slot in the run
    // buffer of every rank (peer memory over NVLink), or just its own (n_peers = 1: an all-gather follows).
    // n_peers = 0: P_new by caller index.
    uint32_t s_begin, s_end;
    int n_peers;
    P4<T>* Cp[WTP_MAX_PEERS];
    T a_lo, a_max;
    ForceP<T> force;
    // density classes (graded clouds, variable spacing): this pass answers the points whose spacing at their current
    // position lies in [cls_lo, cls_hi) on an index whose cell size suits that spacing; cls_hi <= 0: everything
    T cls_lo, cls_hi;
    uint64_t rng_key;      // sweep_key(kick_seed, iteration): the random directions of coincident pairs (common.cuh)
    RepelPartial<T>* partials;
};

constexpr int SW_THREADS = 256;
constexpr int SW_WARPS = SW_THREADS / 32;
constexpr int SW_RUN = 8;                        // consecutive entries per warp and trip
template <class T, int KPL> __host__ __device__ constexpr int sw_tile_cap() { return KPL != 1 ? 0 : (sizeof(T) == 8 ? 224 : 288); }
template <class T> __host__ __device__ constexpr int sw_min_blocks() { return sizeof(T) == 8 ? 3 : 4; }

// the adaptive step of one point (src/repel.jl:282-291): |F| s, alpha_i = clamp(1/(|F| + 1e-30), alpha_lo, alpha_max),
// disp = s alpha_i F capped at one spacing. Returns |F| s; (F0, F1, F2) become the displacement.
template <class T, int D>
__device__ __forceinline__ T step_from_force(T s, T a_lo, T a_max, T& F0, T& F1, T& F2) {
    T n2 = F0 * F0 + F1 * F1;
    if (D == 3) n2 = n2 + F2 * F2;
    const T Fn = fast_sqrt<T>(n2);                                                 // :282
    T ai = fast_div<T>((T)1, Fn + (T)1.0e-30);                                     // :285
    ai = ai > a_max ? a_max : (ai < a_lo ? a_lo : ai);
    const T sa = s * ai;
    F0 = sa * F0; F1 = sa * F1; F2 = D == 3 ? sa * F2 : (T)0;                      // :286
    T dn2 = F0 * F0 + F1 * F1;
    if (D == 3) dn2 = dn2 + F2 * F2;
    if (dn2 > s * s) { const T sc = fast_div<T>(s, fast_sqrt<T>(dn2)); F0 = F0 * sc; F1 = F1 * sc; F2 = F2 * sc; }   // :288-290
    return Fn * s;                                                                 // :283
}

template <class T, int D, int KPL, int FK>
__global__ void __launch_bounds__(SW_THREADS, sw_min_blocks<T>()) repel_sweep_kernel(const SweepArgs<T> a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int CAP = sw_tile_cap<T, KPL>();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t s_bar[SW_WARPS];
    __shared__ T s_term[SW_WARPS][32 * KPL][3];
    __shared__ RepelPartial<T> s_part[SW_WARPS];
    __shared__ __align__(16) unsigned char s_buf[SW_WARPS * 64 * sizeof(Key<T>)];
    P4<T>* tile = reinterpret_cast<P4<T>*>(smem_raw) + (size_t)warp * CAP;
    WarpKnn<T, D, KPL, CAP> knn(a.g, a.sorted, a.cell_start, tile, reinterpret_cast<Key<T>*>(s_buf) + warp * 64, &s_bar[warp], lane);
    RepelPartial<T> acc;
    partial_init(acc);
    const T u0sq = a.force.u0 * a.force.u0;

    // every warp takes runs of SW_RUN consecutive entries (sorted positions, or entries of qlist), dealt round-robin
    const uint32_t nq = a.nq_dev ? *a.nq_dev : a.nq;
    for (uint32_t first = (blockIdx.x * SW_WARPS + warp) * SW_RUN; first < nq; first += gridDim.x * SW_WARPS * SW_RUN)
      for (uint32_t qi = first; qi < min(nq, first + SW_RUN); ++qi) {
        const uint32_t j = a.qlist ? a.qlist[qi] : qi;
        const uint32_t self = idx_of(load_p4<T>(a.sorted + j));
        if (self < a.n_fixed + a.id_lo || self >= a.n_fixed + a.id_hi) continue;   // fixed wall / other rank's point
        const uint32_t id = self - a.n_fixed;
        const T xi0 = a.P_old[(size_t)id * D + 0], xi1 = a.P_old[(size_t)id * D + 1];
        const T xi2 = D == 3 ? a.P_old[(size_t)id * D + (D - 1)] : (T)0;
        knn.run(xi0, xi1, xi2, a.kk);                                             // :259
        const T s = a.s_cur ? a.s_cur[id] : a.s_const;                            // :260
        const T inv_s = fast_div<T>((T)1, s), inv_s2 = inv_s * inv_s;

        bool found = false;
        T nn_d2 = (T)0; uint32_t nn_idx = 0xffffffffu;
#pragma unroll
        for (int e = 0; e < KPL; ++e) {                                           // :270-280
            const int r = e * 32 + lane;
            const uint32_t nj = knn.list.e[e].idx();
            const bool valid = r < a.kk && nj != 0xffffffffu && nj != self;       // skip self BY INDEX (:271)
            T t0 = (T)0, t1 = (T)0, t2 = (T)0;
            if (valid) {
                const T d2 = knn.list.e[e].d2();
                if (d2 > (T)0) {
                    const T w = pair_weight<T, FK>(a.force, u0sq, d2, inv_s2);
                    t0 = w * (xi0 - a.S[(size_t)nj * D + 0]);
                    t1 = w * (xi1 - a.S[(size_t)nj * D + 1]);
                    if (D == 3) t2 = w * (xi2 - a.S[(size_t)nj * D + (D - 1)]);
                } else {                                                           // _safe_direction's random branch (:358-364)
                    T dir[3];
                    random_unit<T, D>(a.rng_key, self, nj, dir);
                    const T f0 = force_at_zero<T, FK>(a.force, u0sq);
                    t0 = f0 * dir[0]; t1 = f0 * dir[1]; t2 = f0 * dir[2];
                }
            }
            s_term[warp][r][0] = t0; s_term[warp][r][1] = t1; s_term[warp][r][2] = t2;
            const unsigned m = __ballot_sync(FULL, valid);
            if (!found && m) {                                                     // first non-self hit (:272-275)
                const int src = __ffs(m) - 1;
                nn_d2 = __shfl_sync(FULL, knn.list.e[e].d2(), src);
                nn_idx = __shfl_sync(FULL, nj, src);
                found = true;
            }
        }
        __syncwarp();
        // F += term_j in ascending (d2, index) order: lanes 0..D-1 each add up one component
        T Fc = (T)0;
        if (lane < D) for (int r = 0; r < a.kk; ++r) Fc = Fc + s_term[warp][r][lane];
        __syncwarp();
        T F0 = __shfl_sync(FULL, Fc, 0), F1 = __shfl_sync(FULL, Fc, 1), F2 = D == 3 ? __shfl_sync(FULL, Fc, 2) : (T)0;
        const T fs = step_from_force<T, D>(s, a.a_lo, a.a_max, F0, F1, F2);
        const T p0 = xi0 + F0, p1 = xi1 + F1, p2 = xi2 + F2;                       // :291 (identity wall)
        if (lane == 0) {
            if (a.n_peers > 0) {
                P4<T> rec; rec.x = p0; rec.y = p1; rec.z = p2; rec.w = idx_bits((T)0, self);
                for (int r = 0; r < a.n_peers; ++r) a.Cp[r][j - a.s_begin] = rec;
            } else {
                a.P_new[(size_t)id * D + 0] = p0;
                a.P_new[(size_t)id * D + 1] = p1;
                if (D == 3) a.P_new[(size_t)id * D + (D - 1)] = p2;
            }
        }
        if (knn.missed) acc.missed += 1;
        const T nn = found ? sqrt(nn_d2) : t_max<T>();
        const T s_mon = a.s_cur ? a.spacings[self] : a.s_const;   // spacing as of the last rebuild (:251, :379); a constant spacing needs no gather
        const T u = nn / s_mon;
        acc.s1 += (double)u; acc.s2 += (double)(u * u); acc.n += 1;
        acc.max_force = fs > acc.max_force ? fs : acc.max_force;
        if (nn < acc.min_nn || (nn == acc.min_nn && id < acc.min_id)) { acc.min_nn = nn; acc.min_id = id; acc.min_nn_idx = nn_idx; }
        acc.lo[0] = p0 < acc.lo[0] ? p0 : acc.lo[0]; acc.hi[0] = p0 > acc.hi[0] ? p0 : acc.hi[0];
        acc.lo[1] = p1 < acc.lo[1] ? p1 : acc.lo[1]; acc.hi[1] = p1 > acc.hi[1] ? p1 : acc.hi[1];
        if (D == 3) { acc.lo[2] = p2 < acc.lo[2] ? p2 : acc.lo[2]; acc.hi[2] = p2 > acc.hi[2] ? p2 : acc.hi[2]; }
      }
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        RepelPartial<T> tot = s_part[0];
        for (int w = 1; w < SW_WARPS; ++w) partial_merge(tot, s_part[w]);
        a.partials[blockIdx.x] = tot;
    }
}

// Rows of D coordinates to scattered caller slots, written by the warp together: the 32 x D words are dealt to the lanes
// so that the D words of a row sit in adjacent lanes and leave as one memory transaction, instead of D stores per lane
// that each touch 32 different rows. Every lane of the warp must call it (valid = false: nothing to store).
template <class T, int D>
__device__ __forceinline__ void warp_store_rows(bool valid, size_t id, T x, T y, T z, T* __restrict__ P, int lane) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
        const int flat = lane + 32 * i, src = flat / D, comp = flat % D;
        const T sx = __shfl_sync(FULL, x, src), sy = __shfl_sync(FULL, y, src), sz = D == 3 ? __shfl_sync(FULL, z, src) : (T)0;
        const unsigned long long sid = __shfl_sync(FULL, (unsigned long long)id, src);
        const bool sv = __shfl_sync(FULL, valid ? 1 : 0, src) != 0;
        if (sv) P[(size_t)sid * D + comp] = comp == 0 ? sx : (comp == 1 ? sy : sz);
    }
}

// ------------------------------------------------------------ tiled sweep
// The same iteration on the CTA-tiled search (knn_tile.cuh): one thread per point. After the selection
// the K nearest snapshot points are read back from the staged slab in ascending (d2, index) order, so the
// force terms are added in the reference's serial order without leaving the thread, and the neighbour
// coordinates come from shared memory instead of a gather from the snapshot. Only valid right after a
// rebuild (the sorted records then carry the current positions). Points the fast path cannot settle go
// to the fail list and through repel_sweep_kernel.
template <class T>
__device__ __forceinline__ void partial_warp_reduce(RepelPartial<T>& p) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        RepelPartial<T> b;
        b.s1 = __shfl_down_sync(FULL, p.s1, o); b.s2 = __shfl_down_sync(FULL, p.s2, o); b.n = __shfl_down_sync(FULL, p.n, o);
        b.missed = __shfl_down_sync(FULL, p.missed, o);
        b.max_force = __shfl_down_sync(FULL, p.max_force, o); b.min_nn = __shfl_down_sync(FULL, p.min_nn, o);
        b.min_id = __shfl_down_sync(FULL, p.min_id, o); b.min_nn_idx = __shfl_down_sync(FULL, p.min_nn_idx, o);
#pragma unroll
        for (int d = 0; d < 3; ++d) { b.lo[d] = __shfl_down_sync(FULL, p.lo[d], o); b.hi[d] = __shfl_down_sync(FULL, p.hi[d], o); }
        partial_merge(p, b);   // lanes >= 32 - o fold garbage, never read: lane 0 ends up with the fixed tree over all 32
    }
}

template <class T, int D, int FK>
__global__ void __launch_bounds__(TK_Q, sizeof(T) == 4 ? 5 : 3)
repel_tile_kernel(const SweepArgs<T> a, const TileFails fails) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ typename TileSearch<T, D>::Shared sh;
    __shared__ RepelPartial<T> s_part[TK_WARPS];
    TileSearch<T, D> ts(a.g, a.sorted, a.cell_start, smem_raw, sh);
    const uint32_t j = a.s_begin + blockIdx.x * TK_Q + threadIdx.x;
    ts.init(j, j < a.s_end, a.n_fixed + a.id_lo, a.n_fixed + a.id_hi);   // fixed wall / other rank's points are not swept
    if (a.cls_hi > (T)0 && ts.query) {                                    // another density class answers this point on its own index
        const T sc = a.s_cur[idx_of(ts.q) - a.n_fixed];
        ts.query = sc >= a.cls_lo && sc < a.cls_hi;
    }
    if (a.n_peers > 0 && ts.active && !ts.query)                          // a fixed point of this rank's run: the record as it is
        for (int r = 0; r < a.n_peers; ++r) a.Cp[r][j - a.s_begin] = ts.q;
    RepelPartial<T> acc;
    partial_init(acc);
    const T u0sq = a.force.u0 * a.force.u0;
    while (ts.next_group()) {
        int status = ts.select(a.kk);                                         // :259
        bool moved = false;                                                   // this thread has a new position for P_new
        size_t moved_id = 0;
        T m0 = (T)0, m1 = (T)0, m2 = (T)0;
        if (status == TK_OK) {
            bool ok = true;
            const uint32_t self = idx_of(ts.q), id = self - a.n_fixed;
            const T xi0 = ts.q.x, xi1 = ts.q.y, xi2 = ts.q.z;                 // == P_old[id] right after a rebuild (:246, :257)
            const T s = a.s_cur ? a.s_cur[id] : a.s_const;                    // :260
            const T inv_s = fast_div<T>((T)1, s), inv_s2 = inv_s * inv_s;
            T F0 = (T)0, F1 = (T)0, F2 = (T)0, nn_d2 = (T)0;
            uint32_t nn_idx = 0xffffffffu;
            Key<T> prev = Key<T>::make((T)0, 0u);
#pragma unroll 2
            for (int r = 0; r < a.kk; ++r) {                                  // :270-280, ascending (d2, index)
                const P4<T> p = lds_p4(ts.tile + ts.my[r]);
                const Key<T> key = Key<T>::make(dist2_rn<T, D>(xi0, xi1, xi2, p.x, p.y, p.z), idx_of(p));
                if (r > 0) ok = ok && prev.less(key);
                prev = key;
                if (key.idx() == self) continue;                              // skip self BY INDEX (:271)
                if (nn_idx == 0xffffffffu) { nn_idx = key.idx(); nn_d2 = key.d2(); }   // :272-275
                if (key.d2() > (T)0) {
                    const T w = pair_weight<T, FK>(a.force, u0sq, key.d2(), inv_s2);
                    F0 = F0 + w * (xi0 - p.x);
                    F1 = F1 + w * (xi1 - p.y);
                    if (D == 3) F2 = F2 + w * (xi2 - p.z);
                } else {                                                      // _safe_direction's random branch (:358-364)
                    T dir[3];
                    random_unit<T, D>(a.rng_key, self, key.idx(), dir);
                    const T f0 = force_at_zero<T, FK>(a.force, u0sq);
                    F0 = F0 + f0 * dir[0]; F1 = F1 + f0 * dir[1]; F2 = F2 + f0 * dir[2];
                }
            }
            status = ts.accept(ok, prev);
            if (status == TK_OK) {
                const T fs = step_from_force<T, D>(s, a.a_lo, a.a_max, F0, F1, F2);
                const T p0 = xi0 + F0, p1 = xi1 + F1, p2 = xi2 + F2;          // :291 (the wall rule follows in its own kernel)
                if (a.n_peers > 0) {
                    P4<T> rec; rec.x = p0; rec.y = p1; rec.z = p2; rec.w = idx_bits((T)0, self);
                    for (int r = 0; r < a.n_peers; ++r) a.Cp[r][ts.j - a.s_begin] = rec;
                } else {
                    moved = true; moved_id = id; m0 = p0; m1 = p1; m2 = p2;
                }
                const T nn = nn_idx != 0xffffffffu ? sqrt(nn_d2) : t_max<T>();
                const T u = nn / (a.s_cur ? a.spacings[self] : a.s_const);     // spacing as of the last rebuild (:251, :379); a constant spacing needs no gather
                RepelPartial<T> one;
                one.s1 = (double)u; one.s2 = (double)(u * u); one.n = 1; one.missed = 0; one.max_force = fs;
                one.min_nn = nn; one.min_id = id; one.min_nn_idx = nn_idx;
                one.lo[0] = one.hi[0] = p0; one.lo[1] = one.hi[1] = p1;
                one.lo[2] = D == 3 ? p2 : t_inf<T>(); one.hi[2] = D == 3 ? p2 : -t_inf<T>();
                partial_merge(acc, one);
            }
        }
        if (a.n_peers == 0) warp_store_rows<T, D>(moved, moved_id, m0, m1, m2, a.P_new, ts.lane);
        ts.report(status, fails);
    }
    partial_warp_reduce(acc);
    if (ts.lane == 0) s_part[ts.warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        RepelPartial<T> tot = s_part[0];
        for (int w = 1; w < TK_WARPS; ++w) partial_merge(tot, s_part[w]);
        a.partials[blockIdx.x] = tot;
    }
}

// fold the per-CTA partials in a fixed order (thread t takes partials t, t+256, ...; then
// a fixed tree over the threads) so the stop test is reproducible run to run
// (CTA b of a grid folds the contiguous slice b of the partials, so a large sweep is folded in two launches:
// many slices -> one partial each, then those -> the total; slices and trees are fixed, so is the result)
template <class T>
__global__ void __launch_bounds__(256) repel_finalize_kernel(const RepelPartial<T>* __restrict__ partials, int n, RepelPartial<T>* __restrict__ out) {
    __shared__ RepelPartial<T> s[256];
    RepelPartial<T> acc;
    partial_init(acc);
    const int per = (n + (int)gridDim.x - 1) / (int)gridDim.x;
    const int lo = (int)blockIdx.x * per, hi = min(n, lo + per);
    out += blockIdx.x;
    for (int i = lo + (int)threadIdx.x; i < hi; i += 256) partial_merge(acc, partials[i]);
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { RepelPartial<T> x = s[threadIdx.x]; partial_merge(x, s[threadIdx.x + o]); s[threadIdx.x] = x; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = s[0];
}

// smallest and largest value of an array (the spacing range that decides the density classes): out[0] = min, out[1] = max
// (stage 1: CTA b writes its pair to out[2b], out[2b + 1]; stage 2: one CTA over those pairs, `pairs` = 1)
template <class T>
__global__ void __launch_bounds__(256) minmax_kernel(const T* __restrict__ v, int64_t n, T* __restrict__ out, int pairs) {
    __shared__ T s_lo[8], s_hi[8];
    T lo = t_inf<T>(), hi = -t_inf<T>();
    out += 2 * blockIdx.x;
    if (pairs) {
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const T a = v[2 * i], b = v[2 * i + 1]; lo = a < lo ? a : lo; hi = b > hi ? b : hi; }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) { const T x = v[i]; lo = x < lo ? x : lo; hi = x > hi ? x : hi; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T a = __shfl_xor_sync(FULL, lo, o), b = __shfl_xor_sync(FULL, hi, o);
        lo = a < lo ? a : lo; hi = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { lo = s_lo[w] < lo ? s_lo[w] : lo; hi = s_hi[w] > hi ? s_hi[w] : hi; }
        out[0] = lo; out[1] = hi;
    }
}

template <class T>
__global__ void __launch_bounds__(256) fill_kernel(T* __restrict__ out, int64_t n, T v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}

// After the all-gather of the ranks' runs (slot records per rank, rank r's first cnt_r are live): every movable
// point's new position goes to its caller slot of P_new.
template <class T, int D>
__global__ void __launch_bounds__(256) scatter_runs_kernel(const P4<T>* __restrict__ C_all, uint32_t slot, uint32_t n_all, uint32_t world,
                                                           uint32_t n_fixed, T* __restrict__ P_new) {
    const uint32_t r = blockIdx.y;
    const uint32_t begin = (uint32_t)(((uint64_t)n_all * r) / world), end = (uint32_t)(((uint64_t)n_all * (r + 1)) / world);
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    P4<T> rec; rec.x = rec.y = rec.z = (T)0; rec.w = idx_bits((T)0, 0u);
    const bool live = t < end - begin;
    if (live) rec = load_p4<T>(C_all + (size_t)r * slot + t);
    const uint32_t self = idx_of(rec);
    const bool movable = live && self >= n_fixed;
    warp_store_rows<T, D>(movable, movable ? (size_t)(self - n_fixed) : 0, rec.x, rec.y, rec.z, P_new, threadIdx.x & 31);
}

template <class T, int D, int KPL, int FK>
static void launch_sweep_kpl(wtp_ctx* ctx, const SweepArgs<T>& a, int nblocks) {
    constexpr size_t smem = (size_t)sw_tile_cap<T, KPL>() * sizeof(P4<T>) * SW_WARPS;
    // the attribute belongs to the (function, device) pair: set per launch, a context may sit on any device
    if (smem > 0) WTP_CUDA_CHECK(cudaFuncSetAttribute(repel_sweep_kernel<T, D, KPL, FK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    repel_sweep_kernel<T, D, KPL, FK><<<nblocks, SW_THREADS, smem, ctx->stream>>>(a);
}
template <class T, int D>
static void launch_sweep(wtp_ctx* ctx, const SweepArgs<T>& a, int nblocks) {
    if (a.kk <= 32) {
        if (a.force.kind == WTP_FORCE_CLIPPED) launch_sweep_kpl<T, D, 1, WTP_FORCE_CLIPPED>(ctx, a, nblocks);
        else launch_sweep_kpl<T, D, 1, -1>(ctx, a, nblocks);
    } else if (a.kk <= 64) launch_sweep_kpl<T, D, 2, -1>(ctx, a, nblocks);
    else launch_sweep_kpl<T, D, 4, -1>(ctx, a, nblocks);
    LAUNCH_CHECK(ctx);
}

template <class T, int D>
static void launch_sweep_tiled(wtp_ctx* ctx, const SweepArgs<T>& a, int nblocks, const TileFails& fails) {
    constexpr size_t smem = tk_smem<T>();
    if (a.force.kind == WTP_FORCE_CLIPPED) {
        WTP_CUDA_CHECK(cudaFuncSetAttribute(repel_tile_kernel<T, D, WTP_FORCE_CLIPPED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        repel_tile_kernel<T, D, WTP_FORCE_CLIPPED><<<nblocks, TK_Q, smem, ctx->stream>>>(a, fails);
    } else {
        WTP_CUDA_CHECK(cudaFuncSetAttribute(repel_tile_kernel<T, D, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        repel_tile_kernel<T, D, -1><<<nblocks, TK_Q, smem, ctx->stream>>>(a, fails);
    }
    LAUNCH_CHECK(ctx);
}

// ------------------------------------------------------------ deposition
// _deposit_escaped! (src/repel.jl:483-514) after a sweep of the mesh-wall method. What does not depend on earlier
// deposits of the pass runs on the device, batched over the escaped volume points: their projection onto the mesh,
// the spacing at the landing sites, the k nearest snapshot points of the sites and those neighbours' current
// positions and flags. The occupancy sweep is serial by design (an earlier deposit must be visible to a later
// candidate) and runs on the host in index order over those small arrays; the deposits are then applied on the device.
__global__ void __launch_bounds__(256) deposit_flag_kernel(uint8_t* __restrict__ escaped, const uint8_t* __restrict__ is_bnd, uint32_t n,
                                                           uint32_t* __restrict__ flags) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const bool e = escaped[id] != 0;
    if (e) escaped[id] = 0;                                                       // :493
    flags[id] = (e && !is_bnd[id]) ? 1u : 0u;                                     // :494
}
template <class T>
__global__ void __launch_bounds__(256) deposit_gather_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ pos, uint32_t n,
                                                             const T* __restrict__ P, uint32_t* __restrict__ ids, T* __restrict__ pts) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n || !flags[id]) return;
    const uint32_t e = pos[id];
    ids[e] = id;
    for (int d = 0; d < 3; ++d) pts[(size_t)e * 3 + d] = P[(size_t)id * 3 + d];
}
template <class T>
__global__ void __launch_bounds__(256) deposit_neighbours_kernel(const uint32_t* __restrict__ nbr, uint32_t n, const T* __restrict__ P,
                                                                 const uint8_t* __restrict__ is_bnd, T* __restrict__ npos, uint8_t* __restrict__ nflag) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t j = nbr[t] - 1u;
    for (int d = 0; d < 3; ++d) npos[(size_t)t * 3 + d] = P[(size_t)j * 3 + d];
    nflag[t] = is_bnd[j];
}
template <class T>
__global__ void __launch_bounds__(256) deposit_apply_kernel(const uint32_t* __restrict__ ids, const T* __restrict__ sites, const int64_t* __restrict__ tris,
                                                            uint32_t n, T* __restrict__ P, uint8_t* __restrict__ is_bnd, int64_t* __restrict__ tri_idx) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t id = ids[t];
    for (int d = 0; d < 3; ++d) P[(size_t)id * 3 + d] = sites[(size_t)t * 3 + d];  // :508
    is_bnd[id] = 1;                                                               // :509
    tri_idx[id] = tris[t];                                                        // :510
}

// returns the number of points deposited. P: the movable points after the sweep (n_fixed = 0: movable id = snapshot index).
template <class T>
static int64_t deposit_pass(wtp_ctx* ctx, MeshBuffers& mb, const IndexBuffers& ib, const Grid<T>& g, T* P, int64_t n_move, int kq,
                            const SpacingP<T>& sp, double ratio) {
    cudaStream_t st = ctx->stream;
    const unsigned nb = (unsigned)((n_move + 255) / 256);
    uint32_t* flags = ctx->d_misc.as<uint32_t>((size_t)n_move);
    uint32_t* pos = ctx->d_qlist.as<uint32_t>((size_t)n_move + 1);
    deposit_flag_kernel<<<nb, 256, 0, st>>>(mb.escaped.get<uint8_t>(), mb.is_bnd.get<uint8_t>(), (uint32_t)n_move, flags);
    LAUNCH_CHECK(ctx);
    exclusive_scan_u32(ctx, const_cast<IndexBuffers&>(ib).scan_tmp, flags, pos, n_move);
    uint32_t* h_n = reinterpret_cast<uint32_t*>(static_cast<char*>(ctx->h_pinned) + 2048);
    WTP_CUDA_CHECK(cudaMemcpyAsync(h_n, pos + n_move, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    WTP_CUDA_CHECK(cudaStreamSynchronize(st));
    const size_t nE = *h_n;
    if (nE == 0) return 0;
    // device scratch: ids | candidate points | sites | landing triangles | spacing at the sites | neighbours | their positions, flags
    const size_t nN = nE * (size_t)kq;
    char* base = static_cast<char*>(ctx->d_misc2.reserve(nE * (4 + 3 * sizeof(T) * 2 + 8 + sizeof(T)) + nN * (4 + 3 * sizeof(T) + 1) + 256));
    auto take = [&](size_t bytes) { char* p = base; base += (bytes + 15) & ~(size_t)15; return p; };
    int64_t* d_tri = reinterpret_cast<int64_t*>(take(nE * 8));
    T* d_pts = reinterpret_cast<T*>(take(nE * 3 * sizeof(T)));
    T* d_site = reinterpret_cast<T*>(take(nE * 3 * sizeof(T)));
    T* d_s = reinterpret_cast<T*>(take(nE * sizeof(T)));
    T* d_npos = reinterpret_cast<T*>(take(nN * 3 * sizeof(T)));
    uint32_t* d_ids = reinterpret_cast<uint32_t*>(take(nE * 4));
    uint32_t* d_nbr = reinterpret_cast<uint32_t*>(take(nN * 4));
    uint8_t* d_nflag = reinterpret_cast<uint8_t*>(take(nN));
    deposit_gather_kernel<T><<<nb, 256, 0, st>>>(flags, pos, (uint32_t)n_move, P, d_ids, d_pts);
    LAUNCH_CHECK(ctx);
    mesh_project<T>(ctx, mb, d_pts, (int64_t)nE, d_site, d_tri);                                    // :495
    spacing_eval<T>(ctx, sp, ctx->bvh, d_site, (int64_t)nE, 3, d_s);                                // spacing(site_pt) :501
    knn_points<T>(ctx, ib, g, 3, kq, d_site, (int64_t)nE, d_nbr);                                   // knn(tree, site, kq) :502
    deposit_neighbours_kernel<T><<<(unsigned)((nN + 255) / 256), 256, 0, st>>>(d_nbr, (uint32_t)nN, P, mb.is_bnd.get<uint8_t>(), d_npos, d_nflag);
    LAUNCH_CHECK(ctx);
    std::vector<uint32_t> ids(nE), nbr(nN);
    std::vector<int64_t> tri(nE);
    std::vector<T> site(nE * 3), s_at(nE), npos(nN * 3);
    std::vector<uint8_t> nflag(nN);
    WTP_CUDA_CHECK(cudaMemcpyAsync(ids.data(), d_ids, nE * 4, cudaMemcpyDeviceToHost, st));
    WTP_CUDA_CHECK(cudaMemcpyAsync(tri.data(), d_tri, nE * 8, cudaMemcpyDeviceToHost, st));
    WTP_CUDA_CHECK(cudaMemcpyAsync(site.data(), d_site, nE * 3 * sizeof(T), cudaMemcpyDeviceToHost, st));
    WTP_CUDA_CHECK(cudaMemcpyAsync(s_at.data(), d_s, nE * sizeof(T), cudaMemcpyDeviceToHost, st));
    WTP_CUDA_CHECK(cudaMemcpyAsync(nbr.data(), d_nbr, nN * 4, cudaMemcpyDeviceToHost, st));
    WTP_CUDA_CHECK(cudaMemcpyAsync(npos.data(), d_npos, nN * 3 * sizeof(T), cudaMemcpyDeviceToHost, st));
    WTP_CUDA_CHECK(cudaMemcpyAsync(nflag.data(), d_nflag, nN, cudaMemcpyDeviceToHost, st));
    WTP_CUDA_CHECK(cudaStreamSynchronize(st));
    // the serial sweep, in index order (ids ascend): a neighbour deposited earlier in this pass sits at its site and is a boundary point
    std::vector<uint32_t> dep_of(nE);      // for the candidates deposited so far: slot -> slot in the deposit list
    std::vector<uint32_t> dep_ids; std::vector<T> dep_sites; std::vector<int64_t> dep_tris;
    std::vector<uint8_t> deposited(nE, 0);
    for (size_t e = 0; e < nE; ++e) {
        if (tri[e] == 0) continue;                                                                  // :498
        const double thr = ratio * (double)s_at[e];                                                 // :501
        const T* sx = &site[e * 3];
        bool occupied = false;
        for (int r = 0; r < kq && !occupied; ++r) {                                                 // :503-506
            const uint32_t j = nbr[e * kq + r] - 1u;
            if (j == ids[e]) continue;
            const T* pj = &npos[(e * kq + r) * 3];
            bool bnd = nflag[e * kq + r] != 0;
            // j among this pass's earlier candidates? (ids is sorted)
            const auto it = std::lower_bound(ids.begin(), ids.begin() + (ptrdiff_t)e, j);
            if (it != ids.begin() + (ptrdiff_t)e && *it == j) {
                const size_t ej = (size_t)(it - ids.begin());
                if (deposited[ej]) { bnd = true; pj = &site[ej * 3]; }
            }
            if (!bnd) continue;
            const T dx = pj[0] - sx[0], dy = pj[1] - sx[1], dz = pj[2] - sx[2];
            occupied = (double)std::sqrt((dx * dx + dy * dy) + dz * dz) < thr;
        }
        if (occupied) continue;
        deposited[e] = 1;
        dep_ids.push_back(ids[e]);
        dep_sites.insert(dep_sites.end(), sx, sx + 3);
        dep_tris.push_back(tri[e]);
    }
    const size_t nD = dep_ids.size();
    if (nD == 0) return 0;
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_ids, dep_ids.data(), nD * 4, cudaMemcpyHostToDevice, st));
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_site, dep_sites.data(), nD * 3 * sizeof(T), cudaMemcpyHostToDevice, st));
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_tri, dep_tris.data(), nD * 8, cudaMemcpyHostToDevice, st));
    deposit_apply_kernel<T><<<(unsigned)((nD + 255) / 256), 256, 0, st>>>(d_ids, d_site, d_tri, (uint32_t)nD, P, mb.is_bnd.get<uint8_t>(), mb.tri_idx.get<int64_t>());
    LAUNCH_CHECK(ctx);
    WTP_CUDA_CHECK(cudaStreamSynchronize(st));                                                      // the host vectors go out of scope
    return (int64_t)nD;
}

// ------------------------------------------------------------- host driver
template <class T>
void relax_device(wtp_ctx* ctx, T* d_snap, int64_t n_fixed, int64_t n_move, int D, const wtp_spacing* sp_in, const T* d_bnd,
                  const wtp_force* fm, const wtp_repel_params* prm, MeshBuffers* mesh, T* conv, wtp_trace_entry* trace,
                  wtp_repel_result* res) {
    WTP_REQUIRE(D == 2 || D == 3, WTP_ERR_BAD_ARG, "D must be 2 or 3");
    WTP_REQUIRE(n_fixed >= 0 && n_move >= 0 && n_fixed + n_move > 0, WTP_ERR_BAD_ARG, "empty snapshot");
    WTP_REQUIRE(prm->rebuild_every >= 1, WTP_ERR_BAD_ARG, "rebuild_every must be >= 1");          // src/repel.jl:74
    WTP_REQUIRE(prm->k >= 1 && prm->max_iters >= 0, WTP_ERR_BAD_ARG, "k must be >= 1 and max_iters >= 0");
    WTP_REQUIRE(prm->kick_after >= 0, WTP_ERR_BAD_ARG, "kick_after must be >= 0");
    WTP_REQUIRE(prm->wall == WTP_WALL_IDENTITY || prm->wall == WTP_WALL_MESH, WTP_ERR_UNSUPPORTED, "user-defined constrain closure cannot cross the C ABI");
    WTP_REQUIRE((prm->wall == WTP_WALL_MESH) == (mesh != nullptr), WTP_ERR_BAD_ARG, "params.wall and the wall mesh argument disagree");
    WTP_REQUIRE(!mesh || D == 3, WTP_ERR_BAD_ARG, "the mesh wall rule is 3-D only (src/repel.jl:123)");
    WTP_REQUIRE(prm->deposit_ratio >= 0, WTP_ERR_BAD_ARG, "deposit_ratio must be >= 0");            // src/repel.jl:143
    WTP_REQUIRE(!(prm->deposit_ratio > 0) || (mesh && n_fixed == 0), WTP_ERR_BAD_ARG, "deposit_ratio belongs to the mesh-wall method (every point movable, src/repel.jl:161-172)");
    WTP_REQUIRE(!(prm->deposit_ratio > 0) || ctx->world == 1, WTP_ERR_UNSUPPORTED, "deposition runs on an unsharded context");
    WTP_REQUIRE(fm->kind >= WTP_FORCE_INVERSE && fm->kind <= WTP_FORCE_STRONG, WTP_ERR_UNSUPPORTED, "user-defined RepelForceModel cannot cross the C ABI");
    WTP_REQUIRE(sp_in->kind >= WTP_SPACING_CONSTANT && sp_in->kind <= WTP_SPACING_BOUNDARY_LAYER, WTP_ERR_UNSUPPORTED, "user-defined spacing callable cannot cross the C ABI");
    WTP_REQUIRE(ctx->world == 1 || ctx->nccl_comm, WTP_ERR_STATE, "repel on a sharded context needs the NCCL communicator (wtp_comm_init with a unique id)");
    const int64_t n_all = n_fixed + n_move;
    WTP_REQUIRE(n_all < (int64_t)0xfffffff0u, WTP_ERR_BAD_ARG, "snapshot too large");
    const int kk = (int)std::min<int64_t>(prm->k, n_all);                                           // :208
    WTP_REQUIRE(kk <= WTP_MAX_K_REPEL, WTP_ERR_K_TOO_LARGE, "k exceeds WTP_MAX_K_REPEL (128 neighbours per sweep)");
    res->iters = 0; res->stop_reason = WTP_STOP_MAX_ITERS; res->last_cv = std::numeric_limits<double>::quiet_NaN();
    if (n_move == 0 || prm->max_iters == 0) {
        // the reference still runs sweeps over zero points and records conv = 0 each time (:293)
        if (n_move == 0) for (int i = 0; i < prm->max_iters; ++i) { conv[i] = (T)0; res->iters = i + 1; if (0.0 < prm->tol) { res->stop_reason = WTP_STOP_TOL; break; } }
        return;
    }
    cudaStream_t st = ctx->stream;
    IndexBuffers& ib = ctx->index[0];
    ctx->last_window_points = 0; ctx->last_window_missed = 0;
    const SpacingP<T> sp{sp_in->kind, (T)sp_in->a, (T)sp_in->b, (T)sp_in->c};
    const ForceP<T> force{fm->kind, (T)fm->beta, (T)fm->u0, (T)fm->gamma};
    const bool variable = sp.kind != WTP_SPACING_CONSTANT;
    if (variable) bvh_build<T>(ctx, ctx->bvh, d_bnd, sp_in->n_bnd, D);

    T* spacings = ctx->d_spacings.as<T>((size_t)n_all);
    T* s_cur = variable ? ctx->d_nn.as<T>((size_t)n_move) : nullptr;
    uint32_t* hint_all = variable ? ctx->d_counts.as<uint32_t>((size_t)n_all) : nullptr;   // per-point BVH start hint (snapshot-global)
    uint32_t* nn_cache = variable ? hint_all + n_fixed : nullptr;
    T* Pa = ctx->d_p_new.as<T>((size_t)2 * n_move * D);
    T* Pb = Pa + (size_t)n_move * D;
    T* S_tail = d_snap + (size_t)n_fixed * D;
    WTP_CUDA_CHECK(cudaMemcpyAsync(Pa, S_tail, (size_t)n_move * D * sizeof(T), cudaMemcpyDeviceToDevice, st));

    // bounding boxes: the fixed wall once, the movable tail from each sweep's partials
    double flo[3] = {0, 0, 0}, fhi[3] = {0, 0, 0}, mlo[3], mhi[3];
    if (n_fixed > 0) compute_bbox<T>(ctx, ib, d_snap, n_fixed, D, flo, fhi);
    compute_bbox<T>(ctx, ib, S_tail, n_move, D, mlo, mhi);

    // spacings = ustrip.(spacing.(snap)) (:209). A variable spacing costs a BVH walk per point: the points are visited
    // in the order of a preliminary index of the snapshot (neighbours in space walk the tree together: 8.4 -> ~3 ms at
    // 2 M points against the caller's order), the walk leaves every point's nearest boundary point behind as the next
    // walk's starting bound, and the first sweep's spacing(xi) (:260) is this same evaluation (same positions): a copy.
    bool s_cur_ready = false;
    if (variable) {
        ScopedPhase ph(ctx->timer, PH_SCAN);
        double lo0[3], hi0[3];
        for (int d = 0; d < 3; ++d) {
            lo0[d] = n_fixed > 0 ? std::min(flo[d], mlo[d]) : mlo[d];
            hi0[d] = n_fixed > 0 ? std::max(fhi[d], mhi[d]) : mhi[d];
        }
        const Grid<T> g0 = make_grid<T>(n_all, D, lo0, hi0, ctx->cell_occupancy, 0.0, kk);
        build_index<T>(ctx, ib, d_snap, n_all, D, g0);
        spacing_eval_ordered<T>(ctx, sp, ctx->bvh, d_snap, n_all, D, spacings, hint_all, false, ib.sorted.get<P4<T>>(), n_all, 0);
        WTP_CUDA_CHECK(cudaMemcpyAsync(s_cur, spacings + n_fixed, (size_t)n_move * sizeof(T), cudaMemcpyDeviceToDevice, st));
        s_cur_ready = true;
    } else {
        spacing_eval<T>(ctx, sp, ctx->bvh, d_snap, n_all, D, spacings);
    }

    const int32_t rank = ctx->rank, world = ctx->world;
    // Sharding. By runs (the default whenever the tiled sweep applies): rank r sweeps the run [gsb, gse) of the
    // spatially sorted order of each iteration's snapshot — the same tiled kernel as on one GPU — writes (position,
    // caller index) records, the runs are all-gathered and scattered into caller order on every rank. By caller
    // range (mesh wall, rebuild_every > 1, k > 32): rank r sweeps the movable ids [id_lo, id_hi) with the general kernel.
    const bool by_runs = world > 1 && kk <= 32 && prm->rebuild_every == 1 && !mesh && std::getenv("WTP_NO_TILED") == nullptr &&
                         std::getenv("WTP_REPEL_BY_RANGE") == nullptr;
    const int64_t id_lo = by_runs ? 0 : wtp_shard_begin(n_move, rank, world), id_hi = by_runs ? n_move : wtp_shard_end(n_move, rank, world);
    const int64_t gsb = by_runs ? wtp_shard_begin(n_all, rank, world) : 0, gse = by_runs ? wtp_shard_end(n_all, rank, world) : n_all;
    const int64_t run_slot = (n_all + world - 1) / world;        // records per rank in the all-gather buffer
    const int nblocks = (int)std::min<int64_t>((n_all + SW_WARPS * SW_RUN - 1) / (SW_WARPS * SW_RUN), (int64_t)kNumSMs * 8);
    // tiled sweep (one thread per point) whenever the list fits one register row and this rank sweeps whole runs
    const bool tiled_ok = kk <= 32 && (world == 1 || by_runs) && std::getenv("WTP_NO_TILED") == nullptr;
    const int n_tiled_blocks = tiled_ok ? (int)((gse - gsb + TK_Q - 1) / TK_Q) : 0;
    // the ranks' runs meet in one buffer per rank: through peer memory (every sweep stores its records into all ranks'
    // buffers, two buffers alternate so that a rank may run ahead of its peers' scatter) or, when the ranks cannot map
    // each other's memory, through an all-gather after the sweep
    const size_t run_bytes = (size_t)run_slot * world * sizeof(P4<T>);
    const bool p2p = by_runs && comm_peer_buffers(ctx, ctx->peers, run_bytes);
    P4<T>* C_local = by_runs && !p2p ? ctx->d_misc2.as<P4<T>>((size_t)run_slot * world) : nullptr;
    uint64_t n_sweeps = 0;
    // Density classes (graded clouds). One cell size serves a range of about 3 in the local spacing h: with fewer than
    // ~1.2 points per cell the 3^D block holds too few candidates, with more than ~35 the slab of a CTA no longer fits
    // its tile. A variable spacing that spans more than that (BoundaryLayerSpacing with bulk / at_wall = 4: a density
    // ratio of 64) leaves a quarter to a third of the points to the general kernel on any single grid. So the spacing
    // range [s_min, s_max] of the snapshot is cut into up to three geometric classes; every class gets its own index of
    // ALL points with a cell size of twice its typical spacing (about 8 points per cell where the class lives), and the
    // tiled sweep runs once per class with the points of the class switched on (by their spacing at their current
    // position). Where a class does not live its CTAs hold no query and skip the staging, so the passes add up to
    // little more than one pass; what is added is an index build per extra class. One GPU, rebuild_every = 1.
    int n_cls = 1;
    T cls_bound[4] = {(T)0, (T)0, (T)0, (T)0};
    double cls_cell[3] = {0, 0, 0};
    if (variable && world == 1 && tiled_ok && prm->rebuild_every == 1 && std::getenv("WTP_NO_CLASSES") == nullptr) {
        const int mm_blocks = (int)std::min<int64_t>((n_all + 255) / 256, (int64_t)kNumSMs * 4);
        T* d_mm = ctx->d_reduce.as<T>(2 + 2 * (size_t)mm_blocks);
        minmax_kernel<T><<<mm_blocks, 256, 0, st>>>(spacings, n_all, d_mm + 2, 0);
        LAUNCH_CHECK(ctx);
        minmax_kernel<T><<<1, 256, 0, st>>>(d_mm + 2, mm_blocks, d_mm, 1);
        LAUNCH_CHECK(ctx);
        T h_mm[2];
        WTP_CUDA_CHECK(cudaMemcpyAsync(h_mm, d_mm, sizeof(h_mm), cudaMemcpyDeviceToHost, st));
        WTP_CUDA_CHECK(cudaStreamSynchronize(st));
        const double s_lo = (double)h_mm[0], s_hi = (double)h_mm[1];
        if (s_lo > 0 && std::isfinite(s_hi) && s_hi / s_lo >= 1.8) {
            n_cls = std::min(3, (int)std::ceil(std::log(s_hi / s_lo) / std::log(1.7)));
            for (int c = 0; c <= n_cls; ++c) cls_bound[c] = (T)(s_lo * std::pow(s_hi / s_lo, (double)c / n_cls));
            for (int c = 0; c < n_cls; ++c) cls_cell[c] = 2.0 * std::sqrt((double)cls_bound[c] * (double)cls_bound[c + 1]);
            cls_bound[0] = (T)0;                       // the end classes are open: spacings move with the points
            cls_bound[n_cls] = t_inf_host<T>();
        }
    }
    // per-CTA partials of the sweep launches of an iteration (per class: general sweep | tiled sweep), folded together
    const size_t partials_per_cls = (size_t)nblocks + (size_t)n_tiled_blocks;
    RepelPartial<T>* partials = ctx->d_reduce.as<RepelPartial<T>>(partials_per_cls * (size_t)n_cls + 256 + 1 + world);
    RepelPartial<T>* d_fold = partials + partials_per_cls * (size_t)n_cls;   // first stage of a two-stage fold
    RepelPartial<T>* d_tot = d_fold + 256;
    RepelPartial<T>* d_all = d_tot + 1;
    RepelPartial<T>* h_tot = static_cast<RepelPartial<T>*>(ctx->h_pinned);
    WTP_REQUIRE(sizeof(RepelPartial<T>) * (size_t)(world + 1) + 64 <= 3072, WTP_ERR_BAD_ARG, "world size too large for the staging buffer");
    uint32_t* h_cnt = reinterpret_cast<uint32_t*>(static_cast<char*>(ctx->h_pinned) + 3072);

    // _maybe_kick! state (pair, r/s, count) and the library's own random stream for the kick direction
    int64_t kick_a = -1, kick_b = -1;
    double kick_rs = 0.0;
    int kick_count = 0;
    std::mt19937_64 kick_rng(prm->kick_seed);
    std::normal_distribution<double> kick_normal(0.0, 1.0);

    Grid<T> g{};
    Grid<T> g_cls[3] = {};
    int passes = 0;
    IndexWindow win;
    bool windowed = false;
    // Whether to window the index is decided once for all ranks: ctx->window_off is per-rank state (a sharded k-NN call
    // on a graded cloud sets it on the ranks whose searches left their window), and ranks that disagreed would repeat
    // different iterations below. Every rank adopts the OR of the ranks' flags; from here on the flag only changes on
    // the merged partials, which are identical on every rank.
    bool window_off = ctx->window_off;
    if (by_runs) {
        uint32_t* d_flag = reinterpret_cast<uint32_t*>(d_all);          // scratch, rewritten by the first sweep's all-gather
        uint32_t* h_flag = reinterpret_cast<uint32_t*>(static_cast<char*>(ctx->h_pinned) + 3072);
        h_flag[0] = window_off ? 1u : 0u;
        WTP_CUDA_CHECK(cudaMemcpyAsync(d_flag + world, h_flag, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        comm_allgather_fixed(ctx, d_flag + world, d_flag, sizeof(uint32_t));
        WTP_CUDA_CHECK(cudaMemcpyAsync(h_flag, d_flag, sizeof(uint32_t) * world, cudaMemcpyDeviceToHost, st));
        WTP_CUDA_CHECK(cudaStreamSynchronize(st));
        for (int r = 0; r < world; ++r) window_off = window_off || h_flag[r] != 0;
    }
    int64_t n_deposited = 0;
    T best_cv_T = t_max<T>();
    int64_t last_impr = 0;
    int it = 1, n_conv = 0;
    const uint32_t* qlist = nullptr;
    uint32_t nq = (uint32_t)n_all;
    while (it <= prm->max_iters) {                                                                   // :243
        const bool rebuild = (it - 1) % prm->rebuild_every == 0;                                     // :245
        if (variable && !(it == 1 && s_cur_ready)) {                                                 // spacing(xi), :260 (and :251)
            ScopedPhase ph(ctx->timer, PH_SCAN);
            // the previous iteration's sorted records give a spatially coherent visiting order, its nearest boundary points the
            // bounds. With density classes the order of the COARSEST class's index: 32 consecutive records of the finest grid
            // are a stick of 30 cells where the cloud is coarse, of the coarsest grid a cell or a few neighbouring ones
            // everywhere, and the warp walks the tree for the union of its lanes (bvh.cu)
            const IndexBuffers& ib_order = ctx->index[n_cls > 1 && it > 1 ? n_cls - 1 : 0];
            spacing_eval_ordered<T>(ctx, sp, ctx->bvh, Pa, n_move, D, s_cur, nn_cache, true, ib_order.sorted.get<P4<T>>(), n_all, n_fixed);
        }
        if (rebuild) {
            WTP_CUDA_CHECK(cudaMemcpyAsync(S_tail, Pa, (size_t)n_move * D * sizeof(T), cudaMemcpyDeviceToDevice, st));   // :246
            if (variable) WTP_CUDA_CHECK(cudaMemcpyAsync(spacings + n_fixed, s_cur, (size_t)n_move * sizeof(T), cudaMemcpyDeviceToDevice, st));   // :251
            double lo[3], hi[3];
            for (int d = 0; d < 3; ++d) {
                lo[d] = n_fixed > 0 ? std::min(flo[d], mlo[d]) : mlo[d];
                hi[d] = n_fixed > 0 ? std::max(fhi[d], mhi[d]) : mhi[d];
            }
            g = n_cls > 1 ? make_grid<T>(n_all, D, lo, hi, 1.0e-6, cls_cell[0], kk) : make_grid<T>(n_all, D, lo, hi, ctx->cell_occupancy, 0.0, kk);
            for (int c = 1; c < n_cls; ++c) {                                                        // the coarser classes' indices
                g_cls[c] = make_grid<T>(n_all, D, lo, hi, 1.0e-6, cls_cell[c], kk);
                build_index<T>(ctx, ctx->index[c], d_snap, n_all, D, g_cls[c]);
            }
            // by runs, constant spacing: only the window of the grid around this rank's run is indexed (grid.cu); the
            // variable spacings visit the points in the order of the whole sorted set, so they keep the whole index
            windowed = by_runs && !variable && !window_off && std::getenv("WTP_NO_WINDOW") == nullptr &&
                       build_index_window<T>(ctx, ib, d_snap, n_all, D, g, gsb, gse, 2, &win, &passes);
            if (!windowed) passes = build_index<T>(ctx, ib, d_snap, n_all, D, g);                    // :252
            if (world > 1 && !by_runs) {
                build_query_list(ctx, ib, n_all, n_fixed + id_lo, n_fixed + id_hi, sizeof(T) == 8, ctx->d_misc, ctx->d_misc2, ctx->d_qlist);
                qlist = ctx->d_qlist.get<uint32_t>();
                nq = (uint32_t)(id_hi - id_lo);
            }
        }
        const int64_t sb = windowed ? gsb - win.P0 : gsb, se = windowed ? gse - win.P0 : gse;   // this rank's run in positions of the index
        SweepArgs<T> a;
        a.g = g; a.sorted = ib.sorted.get<P4<T>>(); a.cell_start = ib.cells();
        a.s_begin = (uint32_t)sb; a.s_end = (uint32_t)se;
        a.n_peers = 0;
        P4<T>* C_all = C_local;                      // where this rank finds every rank's run after the exchange
        if (p2p) {
            const size_t parity = (size_t)(n_sweeps++ & 1u);
            a.n_peers = world;
            for (int r = 0; r < world; ++r)
                a.Cp[r] = reinterpret_cast<P4<T>*>(static_cast<char*>(ctx->peers.base[r]) + parity * ctx->peers.bytes_each) + (size_t)rank * run_slot;
            C_all = reinterpret_cast<P4<T>*>(static_cast<char*>(ctx->peers.base[rank]) + parity * ctx->peers.bytes_each);
        } else if (by_runs) {
            a.n_peers = 1;
            a.Cp[0] = C_local + (size_t)rank * run_slot;
        }
        a.S = d_snap; a.P_old = Pa; a.P_new = Pb; a.s_cur = s_cur; a.spacings = spacings; a.s_const = sp.a;
        a.n_fixed = (uint32_t)n_fixed; a.n_all = (uint32_t)n_all; a.id_lo = (uint32_t)id_lo; a.id_hi = (uint32_t)id_hi;
        a.qlist = qlist; a.nq = nq; a.kk = kk; a.rebuild = rebuild ? 1 : 0;
        a.a_lo = (T)prm->alpha_lo; a.a_max = (T)prm->alpha_max; a.force = force; a.partials = partials;
        a.rng_key = sweep_key(prm->kick_seed, (uint64_t)it);
        a.nq_dev = nullptr;
        a.cls_lo = (T)0; a.cls_hi = (T)0;
        int n_partials = nblocks;
        const bool tiled_now = rebuild && tiled_ok;
        TileFails fails{};
        TileFails fails_cls[3] = {};
        {
            ScopedPhase ph(ctx->timer, PH_QUERY);
            if (tiled_now && n_cls > 1) {
                // one tiled sweep per density class on the class's own index, each followed by the general sweep over what it
                // handed back (sorted positions of that index)
                n_partials = 0;
                for (int c = 0; c < n_cls; ++c) {
                    const IndexBuffers& ibc = ctx->index[c];
                    SweepArgs<T> ac = a;
                    ac.g = c == 0 ? g : g_cls[c]; ac.sorted = ibc.sorted.get<P4<T>>(); ac.cell_start = ibc.cells();
                    ac.cls_lo = cls_bound[c]; ac.cls_hi = cls_bound[c + 1];
                    fails_cls[c] = tile_fails(ctx, n_all, c, n_cls);
                    SweepArgs<T> at = ac;
                    at.qlist = nullptr; at.nq = (uint32_t)n_all; at.partials = partials + partials_per_cls * (size_t)c + nblocks;
                    if (D == 2) launch_sweep_tiled<T, 2>(ctx, at, n_tiled_blocks, fails_cls[c]); else launch_sweep_tiled<T, 3>(ctx, at, n_tiled_blocks, fails_cls[c]);
                    ac.qlist = fails_cls[c].list; ac.nq = (uint32_t)n_all; ac.nq_dev = fails_cls[c].counters;
                    ac.partials = partials + partials_per_cls * (size_t)c;
                    ac.cls_hi = (T)0;                                    // the list holds this class's points only
                    if (D == 2) launch_sweep<T, 2>(ctx, ac, nblocks); else launch_sweep<T, 3>(ctx, ac, nblocks);
                }
                n_partials = (int)(partials_per_cls * (size_t)n_cls);
            } else {
                if (tiled_now) {
                    // tiled sweep over every sorted position, then the general sweep over what it handed back
                    fails = tile_fails(ctx, se - sb);
                    SweepArgs<T> at = a;
                    at.qlist = nullptr; at.nq = (uint32_t)(se - sb); at.partials = partials + nblocks;
                    if (n_tiled_blocks > 0) { if (D == 2) launch_sweep_tiled<T, 2>(ctx, at, n_tiled_blocks, fails); else launch_sweep_tiled<T, 3>(ctx, at, n_tiled_blocks, fails); }
                    a.qlist = fails.list; a.nq = (uint32_t)(se - sb); a.nq_dev = fails.counters;
                    n_partials = nblocks + n_tiled_blocks;
                }
                if (D == 2) launch_sweep<T, 2>(ctx, a, nblocks); else launch_sweep<T, 3>(ctx, a, nblocks);
            }
        }
        if (mesh) {                                                                                  // constrain(id, xi, xi + disp), :291, 448-469
            ScopedPhase ph(ctx->timer, PH_SCAN);
            mesh_wall_apply<T>(ctx, *mesh, Pa, Pb, id_lo, id_hi);
        }
        {
            ScopedPhase ph(ctx->timer, PH_REDUCE);
            if (n_partials > 4096) {   // two-stage fold: 256 slices first
                repel_finalize_kernel<T><<<256, 256, 0, st>>>(partials, n_partials, d_fold);
                LAUNCH_CHECK(ctx);
                repel_finalize_kernel<T><<<1, 256, 0, st>>>(d_fold, 256, d_tot);
            } else {
                repel_finalize_kernel<T><<<1, 256, 0, st>>>(partials, n_partials, d_tot);
            }
            LAUNCH_CHECK(ctx);
        }
        RepelPartial<T> tot;
        if (world > 1) {
            ScopedPhase ph(ctx->timer, PH_COMM);
            if (by_runs) {   // every rank's run of (position, caller index) records, then into caller order
                if (p2p) comm_allgather_fixed(ctx, d_tot, d_all, sizeof(RepelPartial<T>));   // also the barrier: every rank's sweep has stored its records here
                else comm_allgather_fixed(ctx, C_all + (size_t)rank * run_slot, C_all, (size_t)run_slot * sizeof(P4<T>));
                const dim3 grid((unsigned)((run_slot + 255) / 256), (unsigned)world);
                if (D == 2) scatter_runs_kernel<T, 2><<<grid, 256, 0, st>>>(C_all, (uint32_t)run_slot, (uint32_t)n_all, (uint32_t)world, (uint32_t)n_fixed, Pb);
                else scatter_runs_kernel<T, 3><<<grid, 256, 0, st>>>(C_all, (uint32_t)run_slot, (uint32_t)n_all, (uint32_t)world, (uint32_t)n_fixed, Pb);
                LAUNCH_CHECK(ctx);
            } else {
                comm_allgather_rows(ctx, Pb, n_move, (size_t)D * sizeof(T));                        // moved positions of every rank
            }
            if (!p2p) comm_allgather_fixed(ctx, d_tot, d_all, sizeof(RepelPartial<T>));
            WTP_CUDA_CHECK(cudaMemcpyAsync(h_tot, d_all, sizeof(RepelPartial<T>) * world, cudaMemcpyDeviceToHost, st));
            WTP_CUDA_CHECK(cudaStreamSynchronize(st));
            tot = h_tot[0];
            for (int r = 1; r < world; ++r) partial_merge(tot, h_tot[r]);
            if (tot.missed > 0) {
                // some rank's search left its window: every rank sees the same merged count (the only input of this
                // decision), drops the windows for good and repeats this iteration on the whole index — a rank that
                // was on the whole index already simply repeats it too (nothing of it has been committed yet)
                window_off = true;
                ctx->window_off = true;
                ctx->last_window_missed += (int64_t)tot.missed;
                continue;
            }
        } else {
            WTP_CUDA_CHECK(cudaMemcpyAsync(h_tot, d_tot, sizeof(RepelPartial<T>), cudaMemcpyDeviceToHost, st));
            if (tiled_now && n_cls > 1) {
                for (int c = 0; c < n_cls; ++c) WTP_CUDA_CHECK(cudaMemcpyAsync(h_cnt + 8 * c, fails_cls[c].counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            } else if (tiled_now) {
                WTP_CUDA_CHECK(cudaMemcpyAsync(h_cnt, fails.counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            }
            WTP_CUDA_CHECK(cudaStreamSynchronize(st));
            tot = h_tot[0];
            if (tiled_now) {
                ctx->last_tile_sparse = ctx->last_tile_dense = ctx->last_tile_other = 0;
                for (int c = 0; c < (n_cls > 1 ? n_cls : 1); ++c) {
                    ctx->last_tile_sparse += h_cnt[8 * c + 1]; ctx->last_tile_dense += h_cnt[8 * c + 2] + h_cnt[8 * c + 4]; ctx->last_tile_other += h_cnt[8 * c + 3];
                    if (std::getenv("WTP_REPEL_DEBUG") && it <= 2)
                        fprintf(stderr, "[wtp repel] it %d class %d/%d [%g, %g) cell %g: leftovers %u = sparse %u, list overflow %u, ties %u, slab too large %u\n", it, c, n_cls,
                                (double)cls_bound[c], (double)cls_bound[c + 1], n_cls > 1 ? cls_cell[c] : (double)g.c, h_cnt[8 * c], h_cnt[8 * c + 1], h_cnt[8 * c + 2], h_cnt[8 * c + 3], h_cnt[8 * c + 4]);
                }
            }
        }
        for (int d = 0; d < 3; ++d) { mlo[d] = d < D ? (double)tot.lo[d] : 0.0; mhi[d] = d < D ? (double)tot.hi[d] : 0.0; }
        conv[n_conv++] = tot.max_force;                                                              // :293
        if (trace || prm->kick_after > 0) {                                                          // :294-304, 396-403
            const int64_t ig = (int64_t)tot.min_id + n_fixed + 1;
            const int64_t jg = tot.min_nn_idx == 0xffffffffu ? 0 : (int64_t)tot.min_nn_idx + 1;
            T sa_, sb_;
            WTP_CUDA_CHECK(cudaMemcpyAsync(&sa_, spacings + (ig - 1), sizeof(T), cudaMemcpyDeviceToHost, st));
            WTP_CUDA_CHECK(cudaMemcpyAsync(&sb_, spacings + (jg > 0 ? jg - 1 : ig - 1), sizeof(T), cudaMemcpyDeviceToHost, st));
            WTP_CUDA_CHECK(cudaStreamSynchronize(st));
            const T s_pair = (sa_ + sb_) / (T)2;
            const int64_t pa = std::min(ig, jg), pb = std::max(ig, jg);
            const double rs = (double)(tot.min_nn / s_pair);
            if (trace) {
                wtp_trace_entry& te = trace[n_conv - 1];
                te.r = (double)tot.min_nn; te.s = (double)s_pair; te.r_over_s = rs;
                te.idx_a = pa; te.idx_b = pb;
            }
            if (prm->kick_after > 0) {                                                               // _maybe_kick!, :415-433
                const bool frozen = pa == kick_a && pb == kick_b && std::fabs(rs - kick_rs) < 1.0e-8;
                const int count = frozen ? kick_count + 1 : 1;
                kick_a = pa; kick_b = pb; kick_rs = rs; kick_count = count;
                if (count >= prm->kick_after) {
                    // one point of the frozen pair (a volume point if there is one, always a movable one) moves by s/10
                    // in a random direction; p is the configuration the sweep just produced (Pb)
                    const int64_t target = pa > prm->n_protected ? pa : (pb > prm->n_protected ? pb : (pa > n_fixed ? pa : pb));
                    if (target > n_fixed) {
                        T s_t, x[3] = {(T)0, (T)0, (T)0};
                        T* slot = Pb + (size_t)(target - 1 - n_fixed) * D;
                        WTP_CUDA_CHECK(cudaMemcpyAsync(&s_t, spacings + (target - 1), sizeof(T), cudaMemcpyDeviceToHost, st));
                        WTP_CUDA_CHECK(cudaMemcpyAsync(x, slot, (size_t)D * sizeof(T), cudaMemcpyDeviceToHost, st));
                        WTP_CUDA_CHECK(cudaStreamSynchronize(st));
                        T dvec[3], nrm = (T)0;
                        for (int d = 0; d < D; ++d) { dvec[d] = (T)kick_normal(kick_rng); nrm += dvec[d] * dvec[d]; }
                        nrm = std::sqrt(nrm);
                        for (int d = 0; d < D; ++d) x[d] = x[d] + (s_t / (T)10) * (dvec[d] / nrm);
                        WTP_CUDA_CHECK(cudaMemcpyAsync(slot, x, (size_t)D * sizeof(T), cudaMemcpyHostToDevice, st));
                        WTP_CUDA_CHECK(cudaStreamSynchronize(st));
                    }
                    kick_count = 0;
                }
            }
        }
        bool stopped = false, keep_old = false;
        if (prm->stall_after > 0 || prm->cv_target > 0) {                                            // :305-327
            const double nn_ = (double)tot.n;
            const T mu = (T)(tot.s1 / nn_);
            const T var = (T)(tot.s2 / nn_) - mu * mu;
            const T cv = std::sqrt(var > (T)0 ? var : (T)0) / mu;
            res->last_cv = (double)cv;
            if (prm->cv_target > 0 && (double)cv <= prm->cv_target) {
                stopped = true; keep_old = true; res->stop_reason = WTP_STOP_CV_TARGET;              // p .= p_old (:314)
            } else if (prm->stall_after > 0) {
                if ((double)cv < (double)best_cv_T * (1 - 1.0e-3)) { best_cv_T = cv; last_impr = it; }
                else if (it - last_impr >= prm->stall_after) { stopped = true; res->stop_reason = WTP_STOP_STALL; }
            }
        }
        if (!keep_old) std::swap(Pa, Pb);  // Pa = p after the sweep
        if (stopped) break;
        if (mesh && prm->deposit_ratio > 0)                                                          // deposit!(p, tree, i), :328
            n_deposited += deposit_pass<T>(ctx, *mesh, ib, g, Pa, n_move, (int)std::min<int64_t>(prm->k, n_move), sp, prm->deposit_ratio);
        if ((double)conv[n_conv - 1] < prm->tol) { res->stop_reason = WTP_STOP_TOL; break; }          // :329-332
        ++it;
    }
    WTP_CUDA_CHECK(cudaMemcpyAsync(S_tail, Pa, (size_t)n_move * D * sizeof(T), cudaMemcpyDeviceToDevice, st));
    if (mesh && world > 1) {   // every rank returns the landing triangles / escape flags of all points
        comm_allgather_rows(ctx, mesh->tri_idx.get<int64_t>(), n_move, sizeof(int64_t));
        comm_allgather_rows(ctx, mesh->escaped.get<uint8_t>(), n_move, sizeof(uint8_t));
    }
    WTP_CUDA_CHECK(cudaStreamSynchronize(st));
    (void)n_deposited;
    res->iters = n_conv;
    ctx->last_timing = wtp_timing{};
    ctx->last_timing.sort_passes = passes;
    ctx->last_timing.query_launches = n_conv;
    ctx->last_timing.n_cells = g.ncells;
    ctx->last_timing.n_leftover_sparse = ctx->last_tile_sparse;
    ctx->last_timing.n_leftover_dense = ctx->last_tile_dense;
    ctx->last_timing.n_leftover_other = ctx->last_tile_other;
    ctx->last_timing.n_window_points = windowed ? win.M : 0;
    ctx->last_timing.n_window_missed = ctx->last_window_missed;
    ctx->last_timing.n_peer_ranks = p2p ? world : 0;
}

template void relax_device<float>(wtp_ctx*, float*, int64_t, int64_t, int, const wtp_spacing*, const float*, const wtp_force*,
                                  const wtp_repel_params*, MeshBuffers*, float*, wtp_trace_entry*, wtp_repel_result*);
template void relax_device<double>(wtp_ctx*, double*, int64_t, int64_t, int, const wtp_spacing*, const double*, const wtp_force*,
                                   const wtp_repel_params*, MeshBuffers*, double*, wtp_trace_entry*, wtp_repel_result*);

template <class T>
void fill_device(wtp_ctx* ctx, T* d_out, int64_t n, T v) {
    if (n <= 0) return;
    fill_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_out, n, v);
    LAUNCH_CHECK(ctx);
}
template void fill_device<float>(wtp_ctx*, float*, int64_t, float);
template void fill_device<double>(wtp_ctx*, double*, int64_t, double);

}  // namespace wtp
