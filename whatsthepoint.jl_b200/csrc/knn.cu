// knn.cu — k-NN topology kernel: _build_knn_neighbors (src/topology.jl:79-84) and
// search/searchdists (src/neighbors.jl:9-21) for every point of the indexed set.
#include "kernels.cuh"
#include "knn_core.cuh"

namespace wtp {

#define LAUNCH_CHECK(ctx)                         \
    do {                                          \
        (ctx)->launches++;                        \
        WTP_CUDA_CHECK(cudaPeekAtLastError());    \
    } while (0)

constexpr int KNN_THREADS = 256;
constexpr int KNN_WARPS = KNN_THREADS / 32;
constexpr int KNN_QPW = 4;                        // queries per warp
constexpr int KNN_QPB = KNN_WARPS * KNN_QPW;      // 32 consecutive sorted queries per CTA

template <class T, int D, int KPL>
__global__ void __launch_bounds__(KNN_THREADS) knn_kernel(const Grid<T> g, const P4<T>* __restrict__ sorted,
                                                          const uint32_t* __restrict__ cell_start,
                                                          const uint32_t* __restrict__ qlist, uint32_t nq, uint32_t q_begin,
                                                          int K1, int drop, int64_t* __restrict__ out_idx,
                                                          T* __restrict__ out_dist, unsigned long long* __restrict__ expanded) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpKnn<T, D, KPL> s(g, sorted, cell_start);
    const int k_out = K1 - drop;
#pragma unroll 1
    for (int it = 0; it < KNN_QPW; ++it) {
        // at any time the CTA's 8 warps work on 8 neighbouring sorted queries (same cells -> L1 hits)
        const uint32_t qi = blockIdx.x * KNN_QPB + it * KNN_WARPS + warp;
        if (qi >= nq) break;
        const uint32_t j = qlist ? qlist[qi] : qi;
        const P4<T> q = load_p4<T>(sorted + j);
        const int rings = s.run(q.x, q.y, q.z, K1, lane);
        if (rings > 1 && lane == 0 && expanded) atomicAdd(expanded, 1ULL);
        const int64_t row = (int64_t)(idx_of(q) - q_begin) * k_out;
#pragma unroll
        for (int e = 0; e < KPL; ++e) {
            const int r = e * 32 + lane;
            if (r >= drop && r < K1) {
                out_idx[row + r - drop] = (int64_t)s.list.idx[e] + 1;
                if (out_dist) out_dist[row + r - drop] = sqrt(s.list.d2[e]);
            }
        }
    }
}

template <class T, int D>
static void launch_knn(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int K1, int drop, const uint32_t* d_qlist,
                       int64_t nq, int64_t q_begin, int64_t* d_out_idx, T* d_out_dist, unsigned long long* d_exp) {
    const unsigned nblocks = (unsigned)((nq + KNN_QPB - 1) / KNN_QPB);
    const P4<T>* sorted = ib.sorted.get<P4<T>>();
    const uint32_t* cs = ib.cell_start.get<uint32_t>();
#define GO(KPL)                                                                                                     \
    knn_kernel<T, D, KPL><<<nblocks, KNN_THREADS, 0, ctx->stream>>>(g, sorted, cs, d_qlist, (uint32_t)nq,           \
                                                                    (uint32_t)q_begin, K1, drop, d_out_idx,          \
                                                                    d_out_dist, d_exp)
    if (K1 <= 32) GO(1);
    else if (K1 <= 64) GO(2);
    else GO(4);
#undef GO
    LAUNCH_CHECK(ctx);
}

template <class T>
void knn_query(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int64_t N, int D, int K1, int drop_first,
               const uint32_t* d_qlist, int64_t n_queries, int64_t q_begin, int64_t* d_out_idx, T* d_out_dist,
               unsigned long long* d_expanded_counter) {
    (void)N;
    WTP_REQUIRE(K1 >= 1 && K1 <= WTP_MAX_K, WTP_ERR_K_TOO_LARGE, "k exceeds WTP_MAX_K (128 list entries)");
    if (n_queries <= 0) return;
    ScopedPhase ph(ctx->timer, PH_QUERY);
    if (D == 2) launch_knn<T, 2>(ctx, ib, g, K1, drop_first, d_qlist, n_queries, q_begin, d_out_idx, d_out_dist, d_expanded_counter);
    else launch_knn<T, 3>(ctx, ib, g, K1, drop_first, d_qlist, n_queries, q_begin, d_out_idx, d_out_dist, d_expanded_counter);
}
template void knn_query<float>(wtp_ctx*, const IndexBuffers&, const Grid<float>&, int64_t, int, int, int, const uint32_t*, int64_t,
                               int64_t, int64_t*, float*, unsigned long long*);
template void knn_query<double>(wtp_ctx*, const IndexBuffers&, const Grid<double>&, int64_t, int, int, int, const uint32_t*, int64_t,
                                int64_t, int64_t*, double*, unsigned long long*);

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
