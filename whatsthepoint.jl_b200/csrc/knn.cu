// knn.cu — k-NN topology kernel: _build_knn_neighbors (src/topology.jl:79-84) and
// search/searchdists (src/neighbors.jl:9-21) for every point of the indexed set.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "kernels.cuh"
#include "knn_core.cuh"
#include "knn_tile.cuh"

namespace wtp {

#define LAUNCH_CHECK(ctx)                         \
    do {                                          \
        (ctx)->launches++;                        \
        WTP_CUDA_CHECK(cudaPeekAtLastError());    \
    } while (0)

constexpr int KNN_THREADS = 256;
constexpr int KNN_WARPS = KNN_THREADS / 32;
constexpr int KNN_RUN = 8;                        // consecutive entries per warp and trip (tile reuse within a cell)
constexpr int KNN_QPB = KNN_WARPS * KNN_RUN;      // entries per CTA and trip
constexpr int KNN_TILE_CAP = 288;                 // records per warp tile: 3^3 cells x ~8 points + 4.9 sigma

template <class T, int KPL>
__host__ __device__ constexpr int knn_tile_cap() { return KPL != 1 ? 0 : (sizeof(T) == 8 ? 224 : KNN_TILE_CAP); }
template <class T> __host__ __device__ constexpr int knn_min_blocks() { return sizeof(T) == 8 ? 3 : 4; }   // long lists (K > 32) use the general path

template <class T, int D, int KPL>
__global__ void __launch_bounds__(KNN_THREADS, knn_min_blocks<T>()) knn_kernel(const Grid<T> g, const P4<T>* __restrict__ sorted,
                                                          const uint32_t* __restrict__ cell_start,
                                                          const uint32_t* __restrict__ qlist, uint32_t nq, const uint32_t* __restrict__ nq_dev,
                                                          const RowMap rows, int K1, int drop, void* __restrict__ out_idx_v, int out32,
                                                          T* __restrict__ out_dist, unsigned long long* __restrict__ expanded,
                                                          const T* __restrict__ ext_q) {
    constexpr int CAP = knn_tile_cap<T, KPL>();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t s_bar[KNN_WARPS];
    __shared__ __align__(16) unsigned char s_buf[KNN_WARPS * 64 * sizeof(Key<T>)];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    P4<T>* tile = reinterpret_cast<P4<T>*>(smem_raw) + (size_t)warp * CAP;
    WarpKnn<T, D, KPL, CAP> s(g, sorted, cell_start, tile, reinterpret_cast<Key<T>*>(s_buf) + warp * 64, &s_bar[warp], lane);
    const int k_out = K1 - drop;
    int64_t* __restrict__ out_idx = static_cast<int64_t*>(out_idx_v);     // ABI layout: int64, 1-based
    uint32_t* __restrict__ out_idx32 = static_cast<uint32_t*>(out_idx_v); // staging layout of the host entry points (widened on the host)
    auto answer = [&](uint32_t j) {
        // the query: record j of the index, or (ext_q) point j of a separate array, whose row is row j
        P4<T> q;
        if (ext_q) { q.x = ext_q[(size_t)j * D]; q.y = ext_q[(size_t)j * D + 1]; q.z = D == 3 ? ext_q[(size_t)j * D + (D - 1)] : (T)0; q.w = idx_bits((T)0, j); }
        else q = load_p4<T>(sorted + j);
        const int rings = s.run(q.x, q.y, q.z, K1);
        if (rings > 1 && lane == 0 && expanded) atomicAdd(expanded, 1ULL);
        if (s.missed && lane == 0 && expanded) atomicAdd(expanded + 1, 1ULL);   // windowed index: the caller repeats the call on the whole index
        const int64_t row = ext_q ? (int64_t)j * k_out : rows.elem(j, idx_of(q), k_out);
#pragma unroll
        for (int e = 0; e < KPL; ++e) {
            const int r = e * 32 + lane;
            if (r >= drop && r < K1) {
                if (out32) out_idx32[row + r - drop] = s.list.e[e].idx() + 1u;
                else out_idx[row + r - drop] = (int64_t)s.list.e[e].idx() + 1;
                if (out_dist) out_dist[row + r - drop] = sqrt(s.list.e[e].d2());
            }
        }
    };
    // every warp takes runs of KNN_RUN consecutive entries (sorted positions, or entries of qlist), dealt round-robin:
    // neighbours in space reuse the staged cell block. nq_dev: length of a device-built list (the tiled pass's leftovers).
    if (nq_dev) nq = *nq_dev;
#pragma unroll 1
    for (uint32_t first = (blockIdx.x * KNN_WARPS + warp) * KNN_RUN; first < nq; first += gridDim.x * KNN_WARPS * KNN_RUN) {
        const uint32_t end = min(nq, first + KNN_RUN);
#pragma unroll 1
        for (uint32_t qi = first; qi < end; ++qi) answer(qlist ? qlist[qi] : qi + rows.pos_base);   // no list: the run [pos_base, pos_base + nq)
    }
}

template <class T, int D, int KPL>
static void launch_knn_kpl(wtp_ctx* ctx, unsigned nblocks, const Grid<T>& g, const P4<T>* sorted, const uint32_t* cs,
                           const uint32_t* d_qlist, int64_t nq, const uint32_t* d_nq, const RowMap& rows, int K1, int drop,
                           void* d_out_idx, int out32, T* d_out_dist, unsigned long long* d_exp, const T* d_ext_q = nullptr) {
    constexpr size_t smem = (size_t)knn_tile_cap<T, KPL>() * sizeof(P4<T>) * KNN_WARPS;
    // the attribute belongs to the (function, device) pair: set per launch, a context may sit on any device
    if (smem > 48 * 1024) WTP_CUDA_CHECK(cudaFuncSetAttribute(knn_kernel<T, D, KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_kernel<T, D, KPL><<<nblocks, KNN_THREADS, smem, ctx->stream>>>(g, sorted, cs, d_qlist, (uint32_t)nq, d_nq, rows,
                                                                       K1, drop, d_out_idx, out32, d_out_dist, d_exp, d_ext_q);
}

template <class T, int D>
static void launch_knn(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int K1, int drop, const uint32_t* d_qlist,
                       int64_t nq, const uint32_t* d_nq, const RowMap& rows, void* d_out_idx, int out32, T* d_out_dist,
                       unsigned long long* d_exp, const T* d_ext_q = nullptr) {
    // a device-side count (the tiled pass's leftovers): a fixed grid strides over however many entries there are
    const unsigned nblocks = (unsigned)std::min<int64_t>((nq + KNN_QPB - 1) / KNN_QPB, d_nq ? (int64_t)kNumSMs * 4 : (int64_t)0x7fffffff);
    const P4<T>* sorted = ib.sorted.get<P4<T>>();
    const uint32_t* cs = ib.cells();
    if (K1 <= 32) launch_knn_kpl<T, D, 1>(ctx, nblocks, g, sorted, cs, d_qlist, nq, d_nq, rows, K1, drop, d_out_idx, out32, d_out_dist, d_exp, d_ext_q);
    else if (K1 <= 64) launch_knn_kpl<T, D, 2>(ctx, nblocks, g, sorted, cs, d_qlist, nq, d_nq, rows, K1, drop, d_out_idx, out32, d_out_dist, d_exp, d_ext_q);
    else if (K1 <= 128) launch_knn_kpl<T, D, 4>(ctx, nblocks, g, sorted, cs, d_qlist, nq, d_nq, rows, K1, drop, d_out_idx, out32, d_out_dist, d_exp, d_ext_q);
    else launch_knn_kpl<T, D, 8>(ctx, nblocks, g, sorted, cs, d_qlist, nq, d_nq, rows, K1, drop, d_out_idx, out32, d_out_dist, d_exp, d_ext_q);
    LAUNCH_CHECK(ctx);
}

template <class T>
void knn_query(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int64_t N, int D, int K1, int drop_first,
               const uint32_t* d_qlist, int64_t n_queries, const RowMap& rows, void* d_out_idx, T* d_out_dist,
               unsigned long long* d_expanded_counter, bool out32) {
    (void)N;
    WTP_REQUIRE(K1 >= 1 && K1 <= WTP_MAX_K, WTP_ERR_K_TOO_LARGE, "k exceeds WTP_MAX_K (256 list entries)");
    if (n_queries <= 0) return;
    ScopedPhase ph(ctx->timer, PH_QUERY);
    if (D == 2) launch_knn<T, 2>(ctx, ib, g, K1, drop_first, d_qlist, n_queries, nullptr, rows, d_out_idx, out32 ? 1 : 0, d_out_dist, d_expanded_counter);
    else launch_knn<T, 3>(ctx, ib, g, K1, drop_first, d_qlist, n_queries, nullptr, rows, d_out_idx, out32 ? 1 : 0, d_out_dist, d_expanded_counter);
}

template void knn_query<float>(wtp_ctx*, const IndexBuffers&, const Grid<float>&, int64_t, int, int, int, const uint32_t*, int64_t,
                               const RowMap&, void*, float*, unsigned long long*, bool);
template void knn_query<double>(wtp_ctx*, const IndexBuffers&, const Grid<double>&, int64_t, int, int, int, const uint32_t*, int64_t,
                                const RowMap&, void*, double*, unsigned long long*, bool);

// The K nearest index points of arbitrary query points (n_q x D, device): 0-based caller indices + 1 as uint32 rows,
// ascending (d2, index). Used by the deposition pass of repel (knn(tree, site, kq), src/repel.jl:502).
template <class T>
void knn_points(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int D, int K, const T* d_q, int64_t n_q, uint32_t* d_out_idx32) {
    WTP_REQUIRE(K >= 1 && K <= WTP_MAX_K, WTP_ERR_K_TOO_LARGE, "k exceeds WTP_MAX_K (256 list entries)");
    if (n_q <= 0) return;
    const RowMap rows{0u, 0u, 1};
    if (D == 2) launch_knn<T, 2>(ctx, ib, g, K, 0, nullptr, n_q, nullptr, rows, d_out_idx32, 1, nullptr, nullptr, d_q);
    else launch_knn<T, 3>(ctx, ib, g, K, 0, nullptr, n_q, nullptr, rows, d_out_idx32, 1, nullptr, nullptr, d_q);
}
template void knn_points<float>(wtp_ctx*, const IndexBuffers&, const Grid<float>&, int, int, const float*, int64_t, uint32_t*);
template void knn_points<double>(wtp_ctx*, const IndexBuffers&, const Grid<double>&, int, int, const double*, int64_t, uint32_t*);

TileFails tile_fails(wtp_ctx* ctx, int64_t n, int slot, int slots) {
    const size_t each = 16 + (size_t)n;
    uint32_t* base = ctx->d_fail.as<uint32_t>(each * (size_t)slots) + each * (size_t)slot;
    WTP_CUDA_CHECK(cudaMemsetAsync(base, 0, 16 * sizeof(uint32_t), ctx->stream));
    return TileFails{base, base + 16};
}

// Tiled front end (knn_tile.cuh) over the sorted positions [s_begin, s_end), then the general kernel over
// whatever it handed back.
template <class T, int D>
static void run_tiled(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int64_t N, int K1, int drop, int64_t s_begin, int64_t s_end,
                      const RowMap& rows, void* d_out_idx, int out32, T* d_out_dist, unsigned long long* d_exp) {
    constexpr size_t smem = tk_smem<T>();
    WTP_CUDA_CHECK(cudaFuncSetAttribute(knn_tile_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device: set per launch
    const TileFails f = tile_fails(ctx, s_end - s_begin);
    const unsigned nblocks = (unsigned)((s_end - s_begin + TK_Q - 1) / TK_Q);
    knn_tile_kernel<T, D><<<nblocks, TK_Q, smem, ctx->stream>>>(g, ib.sorted.get<P4<T>>(), ib.cells(), (uint32_t)s_begin,
                                                               (uint32_t)s_end, rows, K1, drop, d_out_idx, out32, d_out_dist, f);
    LAUNCH_CHECK(ctx);
    launch_knn<T, D>(ctx, ib, g, K1, drop, f.list, s_end - s_begin, f.counters, rows, d_out_idx, out32, d_out_dist, d_exp);
}

template <class T>
void knn_query_tiled(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int64_t N, int D, int K1, int drop_first,
                     int64_t s_begin, int64_t s_end, const RowMap& rows, void* d_out_idx, T* d_out_dist,
                     unsigned long long* d_expanded_counter, bool out32) {
    WTP_REQUIRE(K1 >= 1 && K1 <= 32, WTP_ERR_K_TOO_LARGE, "the tiled k-NN front end holds lists of at most 32 entries");
    if (s_end <= s_begin) return;
    ScopedPhase ph(ctx->timer, PH_QUERY);
    if (D == 2) run_tiled<T, 2>(ctx, ib, g, N, K1, drop_first, s_begin, s_end, rows, d_out_idx, out32 ? 1 : 0, d_out_dist, d_expanded_counter);
    else run_tiled<T, 3>(ctx, ib, g, N, K1, drop_first, s_begin, s_end, rows, d_out_idx, out32 ? 1 : 0, d_out_dist, d_expanded_counter);
}
template void knn_query_tiled<float>(wtp_ctx*, const IndexBuffers&, const Grid<float>&, int64_t, int, int, int, int64_t, int64_t,
                                     const RowMap&, void*, float*, unsigned long long*, bool);
template void knn_query_tiled<double>(wtp_ctx*, const IndexBuffers&, const Grid<double>&, int64_t, int, int, int, int64_t, int64_t,
                                      const RowMap&, void*, double*, unsigned long long*, bool);

// caller indices (1-based) of the sorted positions [s_begin, s_end): the rows a sharded call wrote
template <class T>
__global__ void __launch_bounds__(256) owned_ids_kernel(const P4<T>* __restrict__ sorted, uint32_t s_begin, uint32_t s_end, int64_t* __restrict__ ids) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (s_begin + t < s_end) ids[t] = (int64_t)idx_of(sorted[s_begin + t]) + 1;
}
template <class T>
__global__ void __launch_bounds__(256) owned_ids32_kernel(const P4<T>* __restrict__ sorted, uint32_t s_begin, uint32_t s_end, uint32_t* __restrict__ ids) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (s_begin + t < s_end) ids[t] = idx_of(sorted[s_begin + t]) + 1u;
}
void owned_ids32(wtp_ctx* ctx, const IndexBuffers& ib, bool f64, int64_t s_begin, int64_t s_end, uint32_t* d_ids) {
    if (s_end <= s_begin) return;
    const unsigned nb = (unsigned)((s_end - s_begin + 255) / 256);
    if (f64) owned_ids32_kernel<double><<<nb, 256, 0, ctx->stream>>>(ib.sorted.get<P4<double>>(), (uint32_t)s_begin, (uint32_t)s_end, d_ids);
    else owned_ids32_kernel<float><<<nb, 256, 0, ctx->stream>>>(ib.sorted.get<P4<float>>(), (uint32_t)s_begin, (uint32_t)s_end, d_ids);
    LAUNCH_CHECK(ctx);
}
// Rows of a sharded host call in ascending caller index (the host then writes forward through the caller's table
// instead of at random rows): sort (caller index, row) pairs with the index's own radix sort scratch — dead once the
// index is built — and gather the rows in that order. d_rows_out / d_dist_out: n x k; d_ids_out: caller indices + 1.
template <class T>
__global__ void __launch_bounds__(256) row_keys_kernel(const P4<T>* __restrict__ sorted, uint32_t s_begin, uint32_t n, uint32_t* __restrict__ keys,
                                                       uint32_t* __restrict__ vals) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) { keys[t] = idx_of(sorted[s_begin + t]); vals[t] = t; }
}
template <class U>
__global__ void __launch_bounds__(256) row_gather_kernel(const U* __restrict__ in, const uint32_t* __restrict__ perm, uint64_t n_elems, uint32_t k,
                                                         U* __restrict__ out) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    const uint64_t r = e / k;
    out[e] = in[(uint64_t)perm[r] * k + (e - r * k)];
}
__global__ void __launch_bounds__(256) ids_plus_one_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ ids) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) ids[t] = keys[t] + 1u;
}
template <class T>
void rows_by_caller_index(wtp_ctx* ctx, IndexBuffers& ib, int64_t N, int64_t s_begin, int64_t n, int k, const uint32_t* d_rows, const T* d_dist,
                          uint32_t* d_rows_out, T* d_dist_out, uint32_t* d_ids_out) {
    if (n <= 0) return;
    uint32_t* keys = ib.keys_a.as<uint32_t>((size_t)n);
    uint32_t* vals = ib.vals_a.as<uint32_t>((size_t)n);
    const unsigned nb = (unsigned)((n + 255) / 256);
    row_keys_kernel<T><<<nb, 256, 0, ctx->stream>>>(ib.sorted.get<P4<T>>(), (uint32_t)s_begin, (uint32_t)n, keys, vals);
    LAUNCH_CHECK(ctx);
    int bits = 1;
    while (bits < 32 && ((uint64_t)1 << bits) < (uint64_t)N) ++bits;
    radix_sort_pairs(ctx, ib, n, bits);
    keys = ib.keys_a.get<uint32_t>();
    vals = ib.vals_a.get<uint32_t>();
    const uint64_t ne = (uint64_t)n * (uint64_t)k;
    row_gather_kernel<uint32_t><<<(unsigned)((ne + 255) / 256), 256, 0, ctx->stream>>>(d_rows, vals, ne, (uint32_t)k, d_rows_out);
    LAUNCH_CHECK(ctx);
    if (d_dist) {
        row_gather_kernel<T><<<(unsigned)((ne + 255) / 256), 256, 0, ctx->stream>>>(d_dist, vals, ne, (uint32_t)k, d_dist_out);
        LAUNCH_CHECK(ctx);
    }
    ids_plus_one_kernel<<<nb, 256, 0, ctx->stream>>>(keys, (uint32_t)n, d_ids_out);
    LAUNCH_CHECK(ctx);
}
template void rows_by_caller_index<float>(wtp_ctx*, IndexBuffers&, int64_t, int64_t, int64_t, int, const uint32_t*, const float*, uint32_t*, float*, uint32_t*);
template void rows_by_caller_index<double>(wtp_ctx*, IndexBuffers&, int64_t, int64_t, int64_t, int, const uint32_t*, const double*, uint32_t*, double*, uint32_t*);

void owned_ids(wtp_ctx* ctx, const IndexBuffers& ib, bool f64, int64_t s_begin, int64_t s_end, int64_t* d_ids) {
    if (s_end <= s_begin) return;
    const unsigned nb = (unsigned)((s_end - s_begin + 255) / 256);
    if (f64) owned_ids_kernel<double><<<nb, 256, 0, ctx->stream>>>(ib.sorted.get<P4<double>>(), (uint32_t)s_begin, (uint32_t)s_end, d_ids);
    else owned_ids_kernel<float><<<nb, 256, 0, ctx->stream>>>(ib.sorted.get<P4<float>>(), (uint32_t)s_begin, (uint32_t)s_end, d_ids);
    LAUNCH_CHECK(ctx);
}

// ------------------------------------------------------------ query lists
// Sharded mode: the sorted positions whose original index lies in [q_begin, q_end), in
// sorted order (keeps the spatial coherence of neighbouring warps).
template <class T>
__global__ void __launch_bounds__(256) qflag_kernel(const P4<T>* __restrict__ sorted, uint32_t N, uint32_t qb, uint32_t qe,
                                                    uint32_t* __restrict__ flags) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    uint32_t i = idx_of(sorted[j]);
    flags[j] = (i >= qb && i < qe) ? 1u : 0u;
}
__global__ void __launch_bounds__(256) qcompact_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ pos,
                                                       uint32_t N, uint32_t* __restrict__ qlist) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    if (flags[j]) qlist[pos[j]] = j;
}

void build_query_list(wtp_ctx* ctx, const IndexBuffers& ib, int64_t N, int64_t q_begin, int64_t q_end, bool f64,
                      DevBuf& flags, DevBuf& scan, DevBuf& qlist) {
    uint32_t* d_flags = flags.as<uint32_t>((size_t)N);
    uint32_t* d_pos = scan.as<uint32_t>((size_t)N + 1);
    uint32_t* d_q = qlist.as<uint32_t>((size_t)std::max<int64_t>(q_end - q_begin, 1));
    const unsigned nb = (unsigned)((N + 255) / 256);
    if (f64) qflag_kernel<double><<<nb, 256, 0, ctx->stream>>>(ib.sorted.get<P4<double>>(), (uint32_t)N, (uint32_t)q_begin, (uint32_t)q_end, d_flags);
    else qflag_kernel<float><<<nb, 256, 0, ctx->stream>>>(ib.sorted.get<P4<float>>(), (uint32_t)N, (uint32_t)q_begin, (uint32_t)q_end, d_flags);
    LAUNCH_CHECK(ctx);
    exclusive_scan_u32(ctx, const_cast<IndexBuffers&>(ib).scan_tmp, d_flags, d_pos, N);
    qcompact_kernel<<<nb, 256, 0, ctx->stream>>>(d_flags, d_pos, (uint32_t)N, d_q);
    LAUNCH_CHECK(ctx);
}

}  // namespace wtp
