// radius.cu — RadiusTopology as CSR: _build_radius_neighbors (src/topology.jl:91-97).
// Two passes over the same traversal: count -> exclusive scan (grid.cu) -> fill.
// The grid's cell size is >= r, so the 3^D block around a query holds every hit.
// Rows are written ascending by index (the canonical order), self removed BY INDEX.
#include "kernels.cuh"
#include "knn_core.cuh"
#include "knn_tile.cuh"

namespace wtp {

#define LAUNCH_CHECK(ctx)                         \
    do {                                          \
        (ctx)->launches++;                        \
        WTP_CUDA_CHECK(cudaPeekAtLastError());    \
    } while (0)

constexpr int RAD_THREADS = 256;
constexpr int RAD_WARPS = RAD_THREADS / 32;

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
                                                                   const uint32_t* __restrict__ qlist, uint32_t nq, const uint32_t* __restrict__ nq_dev,
                                                                   uint32_t q_begin, T r2, uint32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (nq_dev) nq = *nq_dev;          // length of a device-built list (the tiled pass's leftovers)
#pragma unroll 1
    for (uint32_t qi = blockIdx.x * RAD_WARPS + warp; qi < nq; qi += gridDim.x * RAD_WARPS) {
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
                                                                  const uint32_t* __restrict__ qlist, uint32_t nq, const uint32_t* __restrict__ nq_dev,
                                                                  uint32_t q_begin, T r2, const int64_t* __restrict__ offsets,
                                                                  uint32_t* __restrict__ scratch, int64_t* __restrict__ indices) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ uint32_t s_hits[RAD_WARPS][32];
    if (nq_dev) nq = *nq_dev;
#pragma unroll 1
    for (uint32_t qi = blockIdx.x * RAD_WARPS + warp; qi < nq; qi += gridDim.x * RAD_WARPS) {
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

// ------------------------------------------------------------- tiled passes
// The CTA-tiled front end of knn_tile.cuh (128 consecutive sorted queries per CTA, their slab of the grid staged
// once by TMA) with one thread per query. The cell size is >= r, so the 3^D block holds every hit.
//   count: the thread sweeps the cells of its block and counts canonical d2 <= r2, self removed by index.
//   fill:  the thread appends the tile slots of its hits to its list in shared memory. Rows of up to 48 entries are
//          sorted by caller index with a sorting network in registers. Longer rows use the order that is already
//          there: the records of a cell are in ascending caller index (the sort is stable), so the hits are 3^D
//          ascending runs; they are appended cell by cell, each run closed by a sentinel, and merged (the smallest
//          head goes to the row, its run advances).
// A slab that does not fit the tile, or a row longer than the list, goes to the general kernels above through
// the pass's fail list.
template <class T, int D>
struct RadTile {
    static constexpr int NROWS = D == 3 ? 9 : 3;
    static constexpr int NRUNS = NROWS * 3;
    static constexpr int CAP = D == 2 ? 1024 : tk_cap<T>();
    static constexpr int LS = D == 3 ? 130 : 110;       // u16 list slots per thread: 65 / 55 words, odd, so the 32 lists of a warp start in 32 banks
    static constexpr int LCAP = LS - NRUNS - 2;         // hits per row; the rest are the runs' sentinels
    static constexpr int MIN_BLOCKS = D == 2 ? (sizeof(T) == 4 ? 5 : 3) : (sizeof(T) == 4 ? 3 : 2);
    static constexpr size_t SMEM_COUNT = (size_t)CAP * sizeof(P4<T>);
    static constexpr size_t SMEM_FILL = SMEM_COUNT + (size_t)LS * 2 * TK_Q;
    using Search = TileSearch<T, D, CAP>;
};

// tile positions of the cells xa .. xa+2 of row r of the thread's block: [cs[c], cs[c + 1]) (empty beyond the grid)
template <class T, int D, class TS>
__device__ __forceinline__ void block_row_cells(const TS& ts, int r, uint32_t (&cs)[4]) {
    const uint32_t base = ts.sh.run_base[r];
    cs[0] = cs[1] = cs[2] = cs[3] = 0;
    if (base == 0xffffffffu) return;
    const uint32_t shift = ts.sh.run_off[r] - ts.sh.run_begin[r];
    const int xa = ts.cx > 0 ? ts.cx - 1 : 0, xb = ts.cx < ts.g.n[0] - 1 ? ts.cx + 1 : ts.g.n[0] - 1;
    const uint32_t* p = ts.cell_start + base + xa;
#pragma unroll
    for (int c = 0; c < 4; ++c) cs[c] = __ldg(p + min(c, xb - xa + 1)) + shift;
}

template <class T, int D>
__global__ void __launch_bounds__(TK_Q, 4)
radius_tile_count_kernel(const Grid<T> g, const P4<T>* __restrict__ sorted, const uint32_t* __restrict__ cell_start, uint32_t s_begin, uint32_t s_end,
                         uint32_t q_begin, uint32_t q_end, T r2, uint32_t* __restrict__ counts, const TileFails fails) {
    using R = RadTile<T, D>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ typename R::Search::Shared sh;
    typename R::Search ts(g, sorted, cell_start, smem_raw, sh);
    const uint32_t j = s_begin + blockIdx.x * TK_Q + threadIdx.x;
    ts.init(j, j < s_end, q_begin, q_end);   // a sharded context answers the caller range [q_begin, q_end): the other records only shape the slab
    while (ts.next_group()) {
        int status = TK_OK;
        if (ts.in_group && ts.query) {
            if (!ts.fits) status = TK_DENSE;
            else {
                const uint32_t self = idx_of(ts.q);
                uint32_t cnt = 0;
#pragma unroll
                for (int r = 0; r < R::NROWS; ++r) {
                    uint32_t cs[4];
                    block_row_cells<T, D>(ts, r, cs);
#pragma unroll 2
                    for (uint32_t t = cs[0]; t < cs[3]; ++t) {
                        const P4<T> p = lds_p4(ts.tile + t);
                        const bool hit = dist2_rn<T, D>(ts.q.x, ts.q.y, ts.q.z, p.x, p.y, p.z) <= r2 && idx_of(p) != self;
                        cnt += hit ? 1u : 0u;
                    }
                }
                counts[self - q_begin] = cnt;
            }
        }
        ts.report(status, fails);
    }
}

// predicated append of a tile slot to a thread's list: if (p) { *(u16*)addr = t; addr += 2; }. No memory clobber: the
// tile reads around it are plain loads of another part of shared memory and may be scheduled across it; the list is
// read back through lds_u16 (volatile as well, so after the appends).
__device__ __forceinline__ void append_u16_if(uint32_t& addr, uint32_t t, bool p) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u16 [%0], %1;\n\t@p add.u32 %0, %0, 2;\n\t}"
                 : "+r"(addr) : "h"((uint16_t)t), "r"((uint32_t)p));
}
template <class T, int D>
__global__ void __launch_bounds__(TK_Q, RadTile<T, D>::MIN_BLOCKS)
radius_tile_fill_kernel(const Grid<T> g, const P4<T>* __restrict__ sorted, const uint32_t* __restrict__ cell_start, uint32_t s_begin, uint32_t s_end,
                        uint32_t q_begin, uint32_t q_end, T r2, const int64_t* __restrict__ offsets, int64_t* __restrict__ indices, const TileFails fails) {
    using R = RadTile<T, D>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ typename R::Search::Shared sh;
    typename R::Search ts(g, sorted, cell_start, smem_raw, sh);
    const uint32_t j = s_begin + blockIdx.x * TK_Q + threadIdx.x;
    ts.init(j, j < s_end, q_begin, q_end);
    const uint32_t list0 = smem_u32(smem_raw + R::SMEM_COUNT) + (uint32_t)threadIdx.x * (uint32_t)(R::LS * 2);
    const uint32_t lim = list0 + (uint32_t)(R::LS - 1) * 2u;   // a row can never be longer than its count: belt and braces
    // the row's place in the CSR arrays: two scattered loads, issued before the slab is staged
    const uint32_t self = idx_of(ts.q);
    int64_t off = 0;
    uint32_t cnt = 0;
    if (ts.active && ts.query) {
        off = __ldg(offsets + (self - q_begin));
        cnt = (uint32_t)(__ldg(offsets + (self - q_begin) + 1) - off);
    }
    int64_t* __restrict__ out = indices + off;
    // caller index of the record in list slot e (sentinel slots: +inf)
    auto key_at = [&](uint32_t addr) {
        const uint32_t t = lds_u16(addr);
        return t == 0xffffu ? 0xffffffffu : idx_of(ts.tile[t]);
    };
    const uint32_t warp_list0 = list0 - (uint32_t)ts.lane * (uint32_t)(R::LS * 2);
    while (ts.next_group()) {
        int status = TK_OK;
        bool parked = false;      // the sorted row is in the thread's list as words: the warp writes it out, a row per store
        if (ts.in_group && ts.query) {
            if (!ts.fits || cnt > (uint32_t)R::LCAP) status = TK_DENSE;
            else if (cnt > 0 && cnt <= 48u) {
                // ---- short row: the hits in sweep order, then a sorting network on the caller indices in registers
                uint32_t addr = list0;
#pragma unroll
                for (int r = 0; r < R::NROWS; ++r) {
                    uint32_t cs[4];
                    block_row_cells<T, D>(ts, r, cs);
#pragma unroll 2
                    for (uint32_t t = cs[0]; t < cs[3]; ++t) {
                        const P4<T> p = lds_p4(ts.tile + t);
                        const bool hit = dist2_rn<T, D>(ts.q.x, ts.q.y, ts.q.z, p.x, p.y, p.z) <= r2 && idx_of(p) != self;
                        append_u16_if(addr, t, hit);
                        addr = min(addr, lim);
                    }
                }
#define CE(i, j) { const uint32_t lo_ = min(k[i], k[j]); k[j] = max(k[i], k[j]); k[i] = lo_; }
                if (cnt <= 32u) {
                    uint32_t k[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) k[e] = (uint32_t)e < cnt ? idx_of(ts.tile[lds_u16(list0 + 2u * e)]) : 0xffffffffu;
#include "sortnet32.inc"
#pragma unroll
                    for (int e = 0; e < 32; ++e) if ((uint32_t)e < cnt) sts_u32(list0 + 4u * e, k[e]);
                } else {
                    uint32_t k[48];
#pragma unroll
                    for (int e = 0; e < 48; ++e) k[e] = (uint32_t)e < cnt ? idx_of(ts.tile[lds_u16(list0 + 2u * e)]) : 0xffffffffu;
#include "sortnet48full.inc"
#pragma unroll
                    for (int e = 0; e < 48; ++e) if ((uint32_t)e < cnt) sts_u32(list0 + 4u * e, k[e]);
                }
#undef CE
                parked = true;
            } else if (cnt > 0) {
                // ---- long row: the hits run by run (one run per cell), each run closed by a sentinel, then a merge
                uint32_t ptr[R::NRUNS];
                uint32_t addr = list0;
#pragma unroll
                for (int r = 0; r < R::NROWS; ++r) {
                    uint32_t cs[4];
                    block_row_cells<T, D>(ts, r, cs);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        ptr[r * 3 + c] = addr;
#pragma unroll 2
                        for (uint32_t t = cs[c]; t < cs[c + 1]; ++t) {
                            const P4<T> p = lds_p4(ts.tile + t);
                            const bool hit = dist2_rn<T, D>(ts.q.x, ts.q.y, ts.q.z, p.x, p.y, p.z) <= r2 && idx_of(p) != self;
                            append_u16_if(addr, t, hit);
                            addr = min(addr, lim);
                        }
                        sts_u16(addr, 0xffffu);
                        addr = min(addr + 2u, lim);
                    }
                }
                // the smallest head is the next entry of the row; its run advances (selects, no branches: the runs'
                // list positions are distinct, so exactly one matches)
                uint32_t h[R::NRUNS];
#pragma unroll
                for (int i = 0; i < R::NRUNS; ++i) h[i] = key_at(ptr[i]);
#pragma unroll 1
                for (uint32_t o = 0; o < cnt; ++o) {
                    uint32_t m = h[0], ma = ptr[0];
#pragma unroll
                    for (int i = 1; i < R::NRUNS; ++i) { const bool lt = h[i] < m; m = lt ? h[i] : m; ma = lt ? ptr[i] : ma; }
                    out[o] = (int64_t)m + 1;
                    const uint32_t na = ma + 2u, nv = key_at(na);
#pragma unroll
                    for (int i = 0; i < R::NRUNS; ++i) { const bool w = ptr[i] == ma; ptr[i] = w ? na : ptr[i]; h[i] = w ? nv : h[i]; }
                }
            }
        }
        __syncwarp();
        unsigned todo = __ballot_sync(FULL, parked);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t cnt_s = __shfl_sync(FULL, cnt, src);
            int64_t* __restrict__ out_s = indices + __shfl_sync(FULL, off, src);
            const uint32_t words = warp_list0 + (uint32_t)src * (uint32_t)(R::LS * 2);
            for (uint32_t e = (uint32_t)ts.lane; e < cnt_s; e += 32u) out_s[e] = (int64_t)lds_u32(words + 4u * e) + 1;
        }
        __syncwarp();                                                                     // the lists are free for the next group
        ts.report(status, fails);
    }
}

template <class T, int D>
static void launch_radius_tile_count(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int64_t n, int64_t q_begin, int64_t q_end, T r2, uint32_t* d_counts,
                                     const TileFails& f) {
    constexpr size_t smem = RadTile<T, D>::SMEM_COUNT;
    WTP_CUDA_CHECK(cudaFuncSetAttribute(radius_tile_count_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device: set per launch
    radius_tile_count_kernel<T, D><<<(unsigned)((n + TK_Q - 1) / TK_Q), TK_Q, smem, ctx->stream>>>(g, ib.sorted.get<P4<T>>(), ib.cells(), 0u, (uint32_t)n,
                                                                                                   (uint32_t)q_begin, (uint32_t)q_end, r2, d_counts, f);
    LAUNCH_CHECK(ctx);
}
template <class T, int D>
static void launch_radius_tile_fill(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int64_t n, int64_t q_begin, int64_t q_end, T r2, const int64_t* d_offsets,
                                    int64_t* d_indices, const TileFails& f) {
    constexpr size_t smem = RadTile<T, D>::SMEM_FILL;
    WTP_CUDA_CHECK(cudaFuncSetAttribute(radius_tile_fill_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device: set per launch
    radius_tile_fill_kernel<T, D><<<(unsigned)((n + TK_Q - 1) / TK_Q), TK_Q, smem, ctx->stream>>>(g, ib.sorted.get<P4<T>>(), ib.cells(), 0u, (uint32_t)n,
                                                                                                  (uint32_t)q_begin, (uint32_t)q_end, r2, d_offsets, d_indices, f);
    LAUNCH_CHECK(ctx);
}

// Tiled pass over every sorted position + the general kernel over what it handed back. A sharded context answers a
// caller range, whose points lie everywhere in the sorted order: it sweeps every tile with the queries of its range
// switched on (TileSearch::init's keep range) — the staging is not shared out, but the pass is the fast one.
static bool radius_tiled_enabled(const wtp_ctx* ctx) {
    (void)ctx;
    return std::getenv("WTP_NO_TILED") == nullptr;
}

template <class T>
void radius_count(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int64_t N, int D, T r, const uint32_t* d_qlist,
                  int64_t n_queries, int64_t q_begin, uint32_t* d_counts) {
    if (n_queries <= 0) return;
    ScopedPhase ph(ctx->timer, PH_QUERY);
    const T r2 = r * r;
    const uint32_t* d_nq = nullptr;
    unsigned nb = (unsigned)((n_queries + RAD_WARPS - 1) / RAD_WARPS);
    if (radius_tiled_enabled(ctx)) {
        const TileFails f = tile_fails(ctx, N);
        if (D == 2) launch_radius_tile_count<T, 2>(ctx, ib, g, N, q_begin, q_begin + n_queries, r2, d_counts, f);
        else launch_radius_tile_count<T, 3>(ctx, ib, g, N, q_begin, q_begin + n_queries, r2, d_counts, f);
        d_qlist = f.list; d_nq = f.counters;
        nb = std::min<unsigned>(nb, (unsigned)kNumSMs * 8u);
    }
    if (D == 2) radius_count_kernel<T, 2><<<nb, RAD_THREADS, 0, ctx->stream>>>(g, ib.sorted.get<P4<T>>(), ib.cells(), d_qlist, (uint32_t)n_queries, d_nq, (uint32_t)q_begin, r2, d_counts);
    else radius_count_kernel<T, 3><<<nb, RAD_THREADS, 0, ctx->stream>>>(g, ib.sorted.get<P4<T>>(), ib.cells(), d_qlist, (uint32_t)n_queries, d_nq, (uint32_t)q_begin, r2, d_counts);
    LAUNCH_CHECK(ctx);
}

template <class T>
void radius_fill(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int64_t N, int D, T r, const uint32_t* d_qlist,
                 int64_t n_queries, int64_t q_begin, const int64_t* d_offsets, int64_t* d_indices) {
    if (n_queries <= 0) return;
    ScopedPhase ph(ctx->timer, PH_QUERY);
    const T r2 = r * r;
    uint32_t* scratch = ctx->d_misc2.get<uint32_t>();
    const uint32_t* d_nq = nullptr;
    unsigned nb = (unsigned)((n_queries + RAD_WARPS - 1) / RAD_WARPS);
    if (radius_tiled_enabled(ctx)) {
        const TileFails f = tile_fails(ctx, N);
        if (D == 2) launch_radius_tile_fill<T, 2>(ctx, ib, g, N, q_begin, q_begin + n_queries, r2, d_offsets, d_indices, f);
        else launch_radius_tile_fill<T, 3>(ctx, ib, g, N, q_begin, q_begin + n_queries, r2, d_offsets, d_indices, f);
        d_qlist = f.list; d_nq = f.counters;
        nb = std::min<unsigned>(nb, (unsigned)kNumSMs * 8u);
    }
    if (D == 2) radius_fill_kernel<T, 2><<<nb, RAD_THREADS, 0, ctx->stream>>>(g, ib.sorted.get<P4<T>>(), ib.cells(), d_qlist, (uint32_t)n_queries, d_nq, (uint32_t)q_begin, r2, d_offsets, scratch, d_indices);
    else radius_fill_kernel<T, 3><<<nb, RAD_THREADS, 0, ctx->stream>>>(g, ib.sorted.get<P4<T>>(), ib.cells(), d_qlist, (uint32_t)n_queries, d_nq, (uint32_t)q_begin, r2, d_offsets, scratch, d_indices);
    LAUNCH_CHECK(ctx);
}

template void radius_count<float>(wtp_ctx*, const IndexBuffers&, const Grid<float>&, int64_t, int, float, const uint32_t*, int64_t, int64_t, uint32_t*);
template void radius_count<double>(wtp_ctx*, const IndexBuffers&, const Grid<double>&, int64_t, int, double, const uint32_t*, int64_t, int64_t, uint32_t*);
template void radius_fill<float>(wtp_ctx*, const IndexBuffers&, const Grid<float>&, int64_t, int, float, const uint32_t*, int64_t, int64_t, const int64_t*, int64_t*);
template void radius_fill<double>(wtp_ctx*, const IndexBuffers&, const Grid<double>&, int64_t, int, double, const uint32_t*, int64_t, int64_t, const int64_t*, int64_t*);

}  // namespace wtp
