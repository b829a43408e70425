// radius.cu — RadiusTopology as CSR: _build_radius_neighbors (src/topology.jl:91-97).
// Two passes over the same traversal: count -> exclusive scan (grid.cu) -> fill.
// The grid's cell size is >= r, so the 3^D block around a query holds every hit.
// Rows are written ascending by index (the canonical order), self removed BY INDEX.
#include "kernels.cuh"
#include "knn_core.cuh"

namespace wtp {

#define LAUNCH_CHECK(ctx)                         \
    do {                                          \
        (ctx)->launches++;                        \
        WTP_CUDA_CHECK(cudaPeekAtLastError());    \
    } while (0)

constexpr int RAD_THREADS = 256;
constexpr int RAD_WARPS = RAD_THREADS / 32;
constexpr int RAD_QPW = 4;
constexpr int RAD_QPB = RAD_WARPS * RAD_QPW;

// Visits every candidate of the 3^D block whose row/cell lower bound does not exceed r2
// and calls f(hit, idx) warp-synchronously (all 32 lanes call f each step).
template <class T, int D, class F>
__device__ __forceinline__ void radius_sweep(const Grid<T>& g, const P4<T>* __restrict__ sorted,
                                             const uint32_t* __restrict__ cell_start, const P4<T>& q, T r2, int lane, F&& f) {
    const int cx = cell_coord(g, q.x, 0), cy = cell_coord(g, q.y, 1), cz = D == 3 ? cell_coord(g, q.z, 2) : 0;
    const uint32_t self = idx_of(q);
    auto face = [&](int d, int j) { return add_rn(g.lo[d], mul_rn((T)j, g.c)); };
    auto gap = [&](int d, T qq, int c0, int o) -> T {
        if (o == 0) return (T)0;
        T gp = o < 0 ? sub_rn(qq, face(d, c0 + o + 1)) : sub_rn(face(d, c0 + o), qq);
        gp = sub_rn(gp, g.slack);
        return gp > (T)0 ? gp : (T)0;
    };
    const int zlo = D == 3 ? -1 : 0, zhi = D == 3 ? 1 : 0;
    for (int dz = zlo; dz <= zhi; ++dz) {
        const int rz = cz + dz;
        if (rz < 0 || rz >= g.n[2]) continue;
        const T gz = D == 3 ? gap(2, q.z, cz, dz) : (T)0;
        for (int dy = -1; dy <= 1; ++dy) {
            const int ry = cy + dy;
            if (ry < 0 || ry >= g.n[1]) continue;
            const T gy = gap(1, q.y, cy, dy);
            T lb = mul_rn(gy, gy);
            if (D == 3) lb = add_rn(lb, mul_rn(gz, gz));
            if (lb > r2) continue;
            int x0 = cx - 1 < 0 ? 0 : cx - 1, x1 = cx + 1 > g.n[0] - 1 ? g.n[0] - 1 : cx + 1;
            const uint32_t row = ((uint32_t)rz * (uint32_t)g.n[1] + (uint32_t)ry) * (uint32_t)g.n[0];
            const uint32_t begin = cell_start[row + x0], end = cell_start[row + x1 + 1];
            for (uint32_t j0 = begin; j0 < end; j0 += 32) {
                const uint32_t j = j0 + lane;
                bool hit = false;
                uint32_t ci = 0;
                if (j < end) {
                    const P4<T> p = load_p4<T>(sorted + j);
                    ci = idx_of(p);
                    hit = (dist2_rn<T, D>(q.x, q.y, q.z, p.x, p.y, p.z) <= r2) && ci != self;
                }
                f(hit, ci);
            }
        }
    }
}

template <class T, int D>
__global__ void __launch_bounds__(RAD_THREADS) radius_count_kernel(const Grid<T> g, const P4<T>* __restrict__ sorted,
                                                                   const uint32_t* __restrict__ cell_start,
                                                                   const uint32_t* __restrict__ qlist, uint32_t nq, uint32_t q_begin,
                                                                   T r2, uint32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int it = 0; it < RAD_QPW; ++it) {
        const uint32_t qi = blockIdx.x * RAD_QPB + it * RAD_WARPS + warp;
        if (qi >= nq) break;
        const uint32_t j = qlist ? qlist[qi] : qi;
        const P4<T> q = load_p4<T>(sorted + j);
        uint32_t cnt = 0;
        radius_sweep<T, D>(g, sorted, cell_start, q, r2, lane, [&](bool hit, uint32_t) { cnt += __popc(__ballot_sync(FULL, hit)); });
        if (lane == 0) counts[idx_of(q) - q_begin] = cnt;
    }
}

// Rows of up to 32 hits are sorted in registers (one bitonic network); longer rows are
// staged unsorted in `scratch` and placed by rank counting.
template <class T, int D>
__global__ void __launch_bounds__(RAD_THREADS) radius_fill_kernel(const Grid<T> g, const P4<T>* __restrict__ sorted,
                                                                  const uint32_t* __restrict__ cell_start,
                                                                  const uint32_t* __restrict__ qlist, uint32_t nq, uint32_t q_begin,
                                                                  T r2, const int64_t* __restrict__ offsets,
                                                                  uint32_t* __restrict__ scratch, int64_t* __restrict__ indices) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ uint32_t s_hits[RAD_WARPS][32];
#pragma unroll 1
    for (int it = 0; it < RAD_QPW; ++it) {
        const uint32_t qi = blockIdx.x * RAD_QPB + it * RAD_WARPS + warp;
        if (qi >= nq) break;
        const uint32_t j = qlist ? qlist[qi] : qi;
        const P4<T> q = load_p4<T>(sorted + j);
        const uint32_t rowi = idx_of(q) - q_begin;
        const int64_t off = offsets[rowi];
        const uint32_t cnt = (uint32_t)(offsets[rowi + 1] - off);
        if (cnt == 0) continue;
        if (cnt <= 32) {
            // hits are compacted into the warp's shared-memory row in traversal order
            uint32_t filled = 0;
            radius_sweep<T, D>(g, sorted, cell_start, q, r2, lane, [&](bool hit, uint32_t ci) {
                const unsigned m = __ballot_sync(FULL, hit);
                if (hit) s_hits[warp][filled + __popc(m & ((1u << lane) - 1))] = ci;
                filled += __popc(m);
            });
            __syncwarp();
            uint32_t mine = (uint32_t)lane < cnt ? s_hits[warp][lane] : 0xffffffffu;
            __syncwarp();
#pragma unroll
            for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
                for (int jj = k >> 1; jj > 0; jj >>= 1) {
                    const uint32_t o = __shfl_xor_sync(FULL, mine, jj);
                    const bool keep_min = ((lane & jj) == 0) == ((lane & k) == 0);
                    mine = keep_min ? (o < mine ? o : mine) : (o > mine ? o : mine);
                }
            }
            if ((uint32_t)lane < cnt) indices[off + lane] = (int64_t)mine + 1;
        } else {
            uint32_t* tmp = scratch + off;
            uint32_t filled = 0;
            radius_sweep<T, D>(g, sorted, cell_start, q, r2, lane, [&](bool hit, uint32_t ci) {
                const unsigned m = __ballot_sync(FULL, hit);
                if (hit) tmp[filled + __popc(m & ((1u << lane) - 1))] = ci;
                filled += __popc(m);
            });
            __syncwarp();
            for (uint32_t a = lane; a < cnt; a += 32) {
                const uint32_t v = tmp[a];
                uint32_t rank = 0;
                for (uint32_t b = 0; b < cnt; ++b) rank += tmp[b] < v ? 1u : 0u;
                indices[off + rank] = (int64_t)v + 1;
            }
        }
    }
}

template <class T>
void radius_count(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int64_t N, int D, T r, const uint32_t* d_qlist,
                  int64_t n_queries, int64_t q_begin, uint32_t* d_counts) {
    (void)N;
    if (n_queries <= 0) return;
    ScopedPhase ph(ctx->timer, PH_QUERY);
    const unsigned nb = (unsigned)((n_queries + RAD_QPB - 1) / RAD_QPB);
    const T r2 = r * r;
    if (D == 2) radius_count_kernel<T, 2><<<nb, RAD_THREADS, 0, ctx->stream>>>(g, ib.sorted.get<P4<T>>(), ib.cell_start.get<uint32_t>(), d_qlist, (uint32_t)n_queries, (uint32_t)q_begin, r2, d_counts);
    else radius_count_kernel<T, 3><<<nb, RAD_THREADS, 0, ctx->stream>>>(g, ib.sorted.get<P4<T>>(), ib.cell_start.get<uint32_t>(), d_qlist, (uint32_t)n_queries, (uint32_t)q_begin, r2, d_counts);
    LAUNCH_CHECK(ctx);
}

template <class T>
void radius_fill(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int64_t N, int D, T r, const uint32_t* d_qlist,
                 int64_t n_queries, int64_t q_begin, const int64_t* d_offsets, int64_t* d_indices) {
    (void)N;
    if (n_queries <= 0) return;
    ScopedPhase ph(ctx->timer, PH_QUERY);
    const unsigned nb = (unsigned)((n_queries + RAD_QPB - 1) / RAD_QPB);
    const T r2 = r * r;
    uint32_t* scratch = ctx->d_misc2.get<uint32_t>();
    if (D == 2) radius_fill_kernel<T, 2><<<nb, RAD_THREADS, 0, ctx->stream>>>(g, ib.sorted.get<P4<T>>(), ib.cell_start.get<uint32_t>(), d_qlist, (uint32_t)n_queries, (uint32_t)q_begin, r2, d_offsets, scratch, d_indices);
    else radius_fill_kernel<T, 3><<<nb, RAD_THREADS, 0, ctx->stream>>>(g, ib.sorted.get<P4<T>>(), ib.cell_start.get<uint32_t>(), d_qlist, (uint32_t)n_queries, (uint32_t)q_begin, r2, d_offsets, scratch, d_indices);
    LAUNCH_CHECK(ctx);
}

template void radius_count<float>(wtp_ctx*, const IndexBuffers&, const Grid<float>&, int64_t, int, float, const uint32_t*, int64_t, int64_t, uint32_t*);
template void radius_count<double>(wtp_ctx*, const IndexBuffers&, const Grid<double>&, int64_t, int, double, const uint32_t*, int64_t, int64_t, uint32_t*);
template void radius_fill<float>(wtp_ctx*, const IndexBuffers&, const Grid<float>&, int64_t, int, float, const uint32_t*, int64_t, int64_t, const int64_t*, int64_t*);
template void radius_fill<double>(wtp_ctx*, const IndexBuffers&, const Grid<double>&, int64_t, int, double, const uint32_t*, int64_t, int64_t, const int64_t*, int64_t*);

}  // namespace wtp
