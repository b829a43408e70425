// knn_tile.cuh — CTA-tiled k-NN front end (lists of up to 32 entries).
//
// A CTA answers TK_Q consecutive sorted queries. Because the sorted order is row-major by
// cell, the queries of one x-row of cells need one slab of the grid: the x-range of their
// cells widened by one cell, in the 3^(D-1) neighbouring rows. The CTA stages that slab once
// with TMA bulk copies (<= 9 contiguous runs of 16/32-byte records, one mbarrier) — ~10 % halo
// instead of the 27x re-staging of a per-cell tile — and then works in two phases:
//   A  thread per query: sweep the three cells of every row that belong to the query's own
//      3^D block (shared-memory broadcast reads, no cross-lane traffic), test
//      d2 <= r0^2 (the density-guided radius of knn_core.cuh) and append the tile slot of
//      every hit to the query's list in shared memory;
//   B  warp per query: the hits are turned into canonical (d2, index) keys, one
//      32-lane bitonic sort (+ merge) yields the sorted list, the K-th entry is checked against
//      the shell of the block, and the row is written with coalesced stores.
// Anything that does not fit this fast path (slab larger than the tile, fewer than K or more
// than TK_LCAP hits, K-th neighbour not provably inside the block) is appended to a fail list
// and answered by the general warp-per-query kernel (knn_core.cuh), so results are exact and
// identical to it: both evaluate the same candidates with the same arithmetic and order keys
// by (d2, index).
#pragma once
#include "knn_core.cuh"

namespace wtp {

constexpr int TK_Q = 128;          // queries (= threads) per CTA
constexpr int TK_WARPS = TK_Q / 32;
constexpr int TK_LCAP = 48;        // hits kept per query
#ifndef TK_FILTER_WIDE
#define TK_FILTER_WIDE 4           // candidates per trip of the filter loop: their tile reads are issued together (1: one at a time)
#endif
// u16 slots per list: an odd number of words, so the 32 lists of a warp start in 32 different banks; LCAP slots + the
// ones an overflowing list can still reach between two clamps of the write position (one clamp per trip)
constexpr int TK_LSTRIDE = TK_FILTER_WIDE > 1 ? 54 : 50;
#ifndef TK_CAP_F32
#define TK_CAP_F32 1792
#endif
#ifndef TK_SWEEP_UNROLL
#define TK_SWEEP_UNROLL 2
#endif
#ifndef TK_RADIUS_SIGMAS
#define TK_RADIUS_SIGMAS 2.1f
#endif
constexpr int kSweepUnroll = TK_SWEEP_UNROLL;
template <class T> __host__ __device__ constexpr int tk_cap() { return sizeof(T) == 4 ? TK_CAP_F32 : 1664; }   // records per CTA tile
template <class T> __host__ __device__ constexpr size_t tk_smem() { return (size_t)tk_cap<T>() * sizeof(P4<T>) + (size_t)TK_LSTRIDE * TK_Q * sizeof(uint16_t); }

// squared distance from the query to the shell of its 3^D block, conservative; +inf when the
// block reaches the grid border on every side, 0 when nothing can be proved
template <class T, int D>
__device__ __forceinline__ T block_shell2(const Grid<T>& g, T qx, T qy, T qz, int cx, int cy, int cz) {
    auto face = [&](int d, int j) { return add_rn(g.lo[d], mul_rn((T)j, g.c)); };
    T shell = t_inf<T>();
    bool open = false;
    if (cx - 1 > 0) { open = true; shell = fmin(shell, sub_rn(sub_rn(qx, face(0, cx - 1)), g.slack)); }
    if (cx + 1 < g.n[0] - 1) { open = true; shell = fmin(shell, sub_rn(sub_rn(face(0, cx + 2), qx), g.slack)); }
    if (cy - 1 > 0) { open = true; shell = fmin(shell, sub_rn(sub_rn(qy, face(1, cy - 1)), g.slack)); }
    if (cy + 1 < g.n[1] - 1) { open = true; shell = fmin(shell, sub_rn(sub_rn(face(1, cy + 2), qy), g.slack)); }
    if (D == 3) {
        if (cz - 1 > 0) { open = true; shell = fmin(shell, sub_rn(sub_rn(qz, face(2, cz - 1)), g.slack)); }
        if (cz + 1 < g.n[2] - 1) { open = true; shell = fmin(shell, sub_rn(sub_rn(face(2, cz + 2), qz), g.slack)); }
    }
    if (!open) return t_inf<T>();
    if (!(shell > (T)0)) return (T)0;
    return mul_rn(shell, shell);
}

template <class T, int D>
__device__ __forceinline__ T prefilter_radius2(const Grid<T>& g, uint32_t block_n, int K) {   // WarpKnn::set_prefilter_radius
    const float target = (float)K + TK_RADIUS_SIGMAS * sqrtf((float)K) + 1.0f;
    const float c = (float)g.c;
    float r2;
    if (D == 3) { const float r3 = target * 27.0f / (4.18879f * (float)block_n); r2 = c * c * cbrtf(r3 * r3); }
    else r2 = c * c * target * 9.0f / (3.14159265f * (float)block_n);
    return (T)r2;
}

// predicated append used by the filter: if (d <= r) { *(u16*)addr = t; addr += 2; }
// TK_APPEND_NOCLOBBER: the asm does not claim to touch memory, so the tile reads of later candidates may be scheduled
// across it (they never alias the lists); the caller then fences once after the sweep, before the lists are read.
#ifdef TK_APPEND_NOCLOBBER
#define TK_APPEND_CLOBBER
#else
#define TK_APPEND_CLOBBER : "memory"
#endif
__device__ __forceinline__ void append_if_le(uint32_t& addr, uint32_t t, float d, float r) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %2, %3;\n\t@p st.shared.u16 [%0], %1;\n\t@p add.u32 %0, %0, 2;\n\t}"
                 : "+r"(addr) : "h"((uint16_t)t), "f"(d), "f"(r) TK_APPEND_CLOBBER);
}
__device__ __forceinline__ void append_if_le(uint32_t& addr, uint32_t t, double d, double r) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.le.f64 p, %2, %3;\n\t@p st.shared.u16 [%0], %1;\n\t@p add.u32 %0, %0, 2;\n\t}"
                 : "+r"(addr) : "h"((uint16_t)t), "d"(d), "d"(r) TK_APPEND_CLOBBER);
}

// The search state of one thread (= one query) of a tiled CTA. Usage (all threads of the CTA together):
//   TileSearch<T, D> ts(...); ts.init(j, active);
//   while (ts.next_group()) {                     // stages the slab of the next x-row of queries
//       bool ok = ts.select(K);                   // phases A + B: ts.my[0..K) = tile slots of the K best, ascending
//       ... read ts.tile[ts.my[r]], r < K, check the order with ts.verify() ...
//       ts.report(fail, ...);
//   }
template <class T, int D, int CAP_ = tk_cap<T>()>
struct TileSearch {
    static constexpr int NROWS = D == 3 ? 9 : 3;
    static constexpr int CAP = CAP_;                // records of the staged slab
    struct Shared {
        uint64_t bar;
        uint32_t rowid[TK_Q];
        int cx[TK_Q];                  // the x cell of every thread's record: where a group is cut when its slab does not fit
        uint32_t next[2], total, tail; // next[p]: first thread of the group formed in pass p (double-buffered); tail: last thread of the current group
        int x0, x1;
        uint32_t run_begin[NROWS], run_off[NROWS], run_base[NROWS];   // base: first cell id of the row (0xffffffff: outside)
    };
    const Grid<T>& g;
    const P4<T>* __restrict__ sorted;
    const uint32_t* __restrict__ cell_start;
    P4<T>* tile;
    uint16_t* my;          // this thread's list: hits during phase A, sorted tile slots after select()
    Shared& sh;
    int tid, lane, warp;
    uint32_t j;            // sorted position of the query
    bool active;           // j is inside the swept range: the thread takes part in forming the groups
    bool query;            // ... and its record is a query (records filtered out by init() still shape the slab)
    bool in_group, fits, last;
    P4<T> q;
    int cx, cy, cz;
    uint32_t rowid, phase, block_n, pass;
    T shell2, r0sq;

    __device__ __forceinline__ TileSearch(const Grid<T>& g_, const P4<T>* s, const uint32_t* cs, unsigned char* smem_dyn, Shared& sh_)
        : g(g_), sorted(s), cell_start(cs), tile(reinterpret_cast<P4<T>*>(smem_dyn)), sh(sh_) {
        tid = threadIdx.x; lane = tid & 31; warp = tid >> 5;
        my = reinterpret_cast<uint16_t*>(smem_dyn + (size_t)CAP * sizeof(P4<T>)) + tid * TK_LSTRIDE;
        phase = 0; pass = 0; last = false; in_group = false; fits = false;
    }
    // qpos: the position whose block is searched (the record's own coordinates unless the caller overrides them)
    // keep_lo/keep_hi: only records whose caller index lies in [keep_lo, keep_hi) are queries
    __device__ __forceinline__ void init(uint32_t j_, bool in_range, uint32_t keep_lo = 0u, uint32_t keep_hi = 0xffffffffu) {
        j = j_; active = in_range; query = false; block_n = 0;
        q.x = q.y = q.z = (T)0; q.w = idx_bits((T)0, 0u);
        cx = cy = cz = 0;
        if (active) {
            q = load_p4<T>(sorted + j);
            const uint32_t i = idx_of(q);
            query = i >= keep_lo && i < keep_hi;
            cx = cell_coord(g, q.x, 0); cy = cell_coord(g, q.y, 1); cz = D == 3 ? cell_coord(g, q.z, 2) : 0;
        }
        rowid = active ? (uint32_t)cz * (uint32_t)g.n[1] + (uint32_t)cy : 0xffffffffu;
        sh.rowid[tid] = rowid;
        sh.cx[tid] = cx;
        if (tid == 0) { sh.next[0] = 0; mbar_init(&sh.bar, 1); fence_mbar_init(); }
        shell2 = block_shell2<T, D>(g, q.x, q.y, q.z, cx, cy, cz);
    }
    // Stages the slab for the next group of queries (same x-row of cells). False when every query is done.
    __device__ __forceinline__ bool next_group() {
        for (;;) {
            if (last) return false;                        // no barrier on the way out: warps retire as they finish
            __syncthreads();                               // previous group is done with the tile; sh.next[pass] is final
            const uint32_t first = sh.next[pass & 1u];
            if (first >= TK_Q || sh.rowid[first] == 0xffffffffu) return false;
            const uint32_t grp_row = sh.rowid[first];
            in_group = active && (uint32_t)tid >= first && rowid == grp_row;
            const bool tail = in_group && (tid == TK_Q - 1 || sh.rowid[tid + 1] != grp_row);
            if ((uint32_t)tid == first) sh.x0 = cx;
            if (tail) { sh.x1 = cx; sh.tail = (uint32_t)tid; sh.next[(pass + 1u) & 1u] = (uint32_t)tid + 1; }   // the other slot: nobody reads it in this pass
            const bool any_query = __syncthreads_or(in_group && query) != 0;
            ++pass;
            if (!any_query) {                                            // nothing to answer in this row: no staging
                const uint32_t nxt = sh.next[pass & 1u];
                last = nxt >= TK_Q || sh.rowid[nxt] == 0xffffffffu;
                in_group = false;
                continue;
            }
            if (warp == 0) {
                // The slab: the rows of the group's 3^(D-1) neighbourhood over the x-range of its cells widened by one. Where
                // the density rises steeply across the rows (graded clouds: a neighbouring row nearer the wall holds several
                // times the points) the slab of a whole group does not fit the tile although the group's own row is sparse:
                // the group is then cut at half its x-range (again and again, down to one column of cells) and the rest of
                // its threads form the next group.
                const int gx0 = sh.x0;
                int gx1 = sh.x1;
                uint32_t begin = 0, len = 0, base = 0xffffffffu, incl = 0, total = 0;
                for (;;) {
                    begin = 0; len = 0; base = 0xffffffffu;
                    if (lane < NROWS) {
                        const int gy = (int)(grp_row % (uint32_t)g.n[1]), gz = (int)(grp_row / (uint32_t)g.n[1]);
                        const int ry = gy + row_dy(lane), rz = D == 3 ? gz + row_dz(lane) : 0;
                        if (ry >= 0 && ry < g.n[1] && rz >= 0 && rz < g.n[2]) {
                            const int x0 = gx0 > 0 ? gx0 - 1 : 0, x1 = gx1 < g.n[0] - 1 ? gx1 + 1 : g.n[0] - 1;
                            base = ((uint32_t)rz * (uint32_t)g.n[1] + (uint32_t)ry) * (uint32_t)g.n[0];
                            begin = cell_start[base + x0];
                            len = cell_start[base + x1 + 1] - begin;
                        }
                    }
                    incl = len;
#pragma unroll
                    for (int o = 1; o < 16; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(FULL, incl, o);
                        if (lane >= o) incl += t;
                    }
                    total = __shfl_sync(FULL, incl, NROWS - 1);
                    if (total <= (uint32_t)CAP || gx1 == gx0) break;
                    gx1 = gx0 + (gx1 - gx0) / 2;
                }
                if (gx1 != sh.x1) {                                      // cut: the last thread of the group whose cell is still inside
                    const uint32_t old_tail = sh.tail;
                    uint32_t best = first;
#pragma unroll
                    for (int jj = 0; jj < TK_Q / 32; ++jj) {
                        const uint32_t t = (uint32_t)lane * (TK_Q / 32) + (uint32_t)jj;
                        if (t >= first && t <= old_tail && sh.cx[t] <= gx1) best = max(best, t);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
                    __syncwarp();
                    if (lane == 0) { sh.x1 = gx1; sh.tail = best; sh.next[pass & 1u] = best + 1; }
                }
                if (lane < NROWS) { sh.run_begin[lane] = begin; sh.run_off[lane] = incl - len; sh.run_base[lane] = base; }
                if (lane == 0) { sh.total = total; if (total > 0 && total <= (uint32_t)CAP) mbar_expect_tx(&sh.bar, total * (uint32_t)sizeof(P4<T>)); }
                __syncwarp();
                if (total > 0 && total <= (uint32_t)CAP && len > 0) tma_bulk_g2s(tile + (incl - len), sorted + begin, len * (uint32_t)sizeof(P4<T>), &sh.bar);
            }
            __syncthreads();
            in_group = in_group && (uint32_t)tid <= sh.tail;
            const uint32_t nxt = sh.next[pass & 1u];
            last = nxt >= TK_Q || sh.rowid[nxt] == 0xffffffffu;
            const uint32_t total = sh.total;
            fits = total <= (uint32_t)CAP;
            if (fits && total > 0) { mbar_wait(&sh.bar, phase); phase ^= 1u; }
            return true;
        }
    }
    // Phases A and B for the query at (x, y, z) in this thread's cell block. On success my[0..K) holds the tile
    // slots of the K nearest candidates in ascending order of their 32-bit key images and the return value is
    // true; the caller still has to confirm the order on full keys (verify) and the K-th key (accept).
    __device__ __forceinline__ int select(int K) {
        if (!in_group || !query) return TK_SPARSE;         // not this thread's turn (never reported)
        // ---- phase A. The filter uses a fused (cheaper) distance and a radius padded by a few ulps, so the
        // list holds at least every block point with canonical d2 <= r0sq.
        uint32_t b[NROWS], e[NROWS];
        block_n = 0;
        const int xa = cx > 0 ? cx - 1 : 0, xb = cx < g.n[0] - 1 ? cx + 1 : g.n[0] - 1;
#pragma unroll
        for (int r = 0; r < NROWS; ++r) {
            const uint32_t base = sh.run_base[r];
            b[r] = e[r] = 0;
            if (base != 0xffffffffu) {
                const uint32_t shift = sh.run_off[r] - sh.run_begin[r];
                b[r] = __ldg(cell_start + base + xa) + shift;
                e[r] = __ldg(cell_start + base + xb + 1) + shift;
                block_n += e[r] - b[r];
            }
        }
        if (!fits) return TK_NOFIT;
        if (block_n < (uint32_t)K) return TK_SPARSE;
        r0sq = prefilter_radius2<T, D>(g, block_n, K);
        const T r0pad = r0sq * ((T)1 + (T)8 * (sizeof(T) == 4 ? (T)1.1920929e-7 : (T)2.220446049250313e-16));
        const uint32_t my_s = smem_u32(my), lim = my_s + (uint32_t)(TK_LCAP + 1) * 2u;
        uint32_t addr = my_s;
#if TK_FILTER_WIDE > 1
        // TK_FILTER_WIDE candidates per trip: their tile reads are independent and issued back to back, so the latency of
        // one shared-memory read is paid once per trip instead of once per candidate. Reads past the end of the run stay
        // inside the CTA's shared memory (the lists follow the tile) and are masked by position.
        const T t_far = t_inf<T>();
#pragma unroll
        for (int r = 0; r < NROWS; ++r) {
#pragma unroll 1
            for (uint32_t t = b[r]; t < e[r]; t += TK_FILTER_WIDE) {
                T d[TK_FILTER_WIDE];
#pragma unroll
                for (int w = 0; w < TK_FILTER_WIDE; ++w) {
                    const P4<T> p = lds_p4(tile + t + w);
                    const T dx = q.x - p.x, dy = q.y - p.y;
                    d[w] = fma(dy, dy, dx * dx);
                    if (D == 3) { const T dz = q.z - p.z; d[w] = fma(dz, dz, d[w]); }
                }
#pragma unroll
                for (int w = 0; w < TK_FILTER_WIDE; ++w) {
                    if (w > 0) d[w] = t + w < e[r] ? d[w] : t_far;
                    append_if_le(addr, t + w, d[w], r0pad);
                }
                addr = min(addr, lim);                     // slots LCAP .. LCAP + WIDE absorb an overflowing list
            }
        }
#else
#pragma unroll
        for (int r = 0; r < NROWS; ++r) {
#pragma unroll kSweepUnroll
            for (uint32_t t = b[r]; t < e[r]; ++t) {
                const P4<T> p = lds_p4(tile + t);
                const T dx = q.x - p.x, dy = q.y - p.y;
                T d = fma(dy, dy, dx * dx);
                if (D == 3) { const T dz = q.z - p.z; d = fma(dz, dz, d); }
                append_if_le(addr, t, d, r0pad);
                addr = min(addr, lim);                     // slots LCAP, LCAP+1 absorb an overflowing list
            }
        }
#endif
        asm volatile("" ::: "memory");                     // the appends are done before the lists are read
        const uint32_t cnt = (addr - my_s) >> 1;
        if (cnt < (uint32_t)K) return TK_SPARSE;
        if (cnt > (uint32_t)TK_LCAP) return TK_DENSE;
        // ---- phase B. Canonical keys of the hits, reduced to 32-bit images (top 26 bits of d2's pattern, the
        // list slot below), go through a fixed sorting network in registers.
        uint32_t k[TK_LCAP];
#pragma unroll
        for (int s = 0; s < TK_LCAP; ++s) {
            const bool live = (uint32_t)s < cnt;
            const uint32_t t = live ? (uint32_t)my[s] : 0u;
            const P4<T> p = lds_p4(tile + t);
            const Key<T> key = Key<T>::make(dist2_rn<T, D>(q.x, q.y, q.z, p.x, p.y, p.z), 0u);
            k[s] = live ? ((key.coarse() & ~63u) | (uint32_t)s) : 0xffffffffu;
        }
#define CE(i, j) { const uint32_t lo_ = min(k[i], k[j]); k[j] = max(k[i], k[j]); k[i] = lo_; }
#include "sortnet48.inc"
#undef CE
        // sorted slots -> sorted tile positions, written back over the head of the list (all reads first), so
        // that the caller's loop over the K best can stay rolled; look-alike mask of neighbouring images
        uint16_t tt[33];
        uint32_t alike = 0;
#pragma unroll
        for (int r = 0; r < 33; ++r) tt[r] = my[min(k[r] & 63u, (uint32_t)TK_LCAP)];   // sentinels (slot 63) stay inside the list
#pragma unroll
        for (int r = 0; r < 32; ++r) alike |= ((k[r] >> 6) == (k[r + 1] >> 6) ? 1u : 0u) << r;
#pragma unroll
        for (int r = 0; r < 33; ++r) my[r] = tt[r];
        return ((alike >> (K - 1)) & 1u) == 0u ? TK_OK : TK_OTHER;   // the first key left out must not be a look-alike of the K-th
    }
    // the K-th key must lie inside the guaranteed radius of the list and inside the block's shell; the order of
    // the K best must have been confirmed on full keys
    __device__ __forceinline__ int accept(bool ordered, const Key<T>& kth) const {
        if (!ordered) return TK_OTHER;
        return kth.d2() <= r0sq && kth.d2() < shell2 ? TK_OK : TK_SPARSE;
    }
    // queries that could not be settled here go to the general kernel's list (warp-aggregated append)
    __device__ __forceinline__ void report(int status, const TileFails& f) const {
        const bool failed_here = in_group && query && status != TK_OK;
        const unsigned m = __ballot_sync(FULL, failed_here);
        if (m == 0) return;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(f.counters, (uint32_t)__popc(m));
        base = __shfl_sync(FULL, base, 0);
        if (failed_here) {
            f.list[base + __popc(m & ((1u << lane) - 1u))] = j;
            atomicAdd(f.counters + status, 1u);
        }
    }
};

template <class T, int D>
__global__ void __launch_bounds__(TK_Q, sizeof(T) == 4 ? 5 : 3)
knn_tile_kernel(const Grid<T> g, const P4<T>* __restrict__ sorted, const uint32_t* __restrict__ cell_start, uint32_t s_begin, uint32_t s_end,
                const RowMap rows, int K1, int drop, void* __restrict__ out_idx_v, int out32, T* __restrict__ out_dist,
                const TileFails fails) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ typename TileSearch<T, D>::Shared sh;
    TileSearch<T, D> ts(g, sorted, cell_start, smem_raw, sh);
    const uint32_t j = s_begin + blockIdx.x * TK_Q + threadIdx.x;
    ts.init(j, j < s_end);
    const int k_out = K1 - drop;
    int64_t* __restrict__ out_idx = static_cast<int64_t*>(out_idx_v);
    uint32_t* __restrict__ out_idx32 = static_cast<uint32_t*>(out_idx_v);
    // Rows leave the CTA a row at a time: each thread parks the caller indices of its K best in its own list (as
    // 32-bit words over the 16-bit slots it has consumed), then the lanes of the warp write one row per store
    // instruction, k contiguous entries. A thread storing its own row entry by entry would touch 32 scattered rows
    // (32 sectors, 32 pages) per store instruction instead of one. Rows with distances (searchdists) and lists longer
    // than the slot area keep the per-thread stores.
    const uint32_t my_s = smem_u32(ts.my);                                              // this thread's list, as 16-bit slots and as K1 words
    const uint32_t warp_s = my_s - (uint32_t)ts.lane * (uint32_t)(TK_LSTRIDE * 2);
    const bool coop = K1 <= TK_LSTRIDE / 2 && out_dist == nullptr;                       // the row fits the list as words
    while (ts.next_group()) {
        int status = ts.select(K1);
        int64_t row = 0;
        if (status == TK_OK) {
            // re-read the K best from the last to the first, check strict canonical order (the images of two keys
            // collide only when their d2 agree to ~17 bits, or on exact ties) and park the indices: word r overwrites
            // slots 2r and 2r+1, both already consumed on the way down
            row = rows.elem(ts.j, idx_of(ts.q), k_out);
            asm volatile("" ::: "memory");                                                // select()'s plain stores to the list come first
            Key<T> next = Key<T>::make((T)0, 0u), kth = next;
            bool ok = true;
#pragma unroll 2
            for (int r = K1 - 1; r >= 0; --r) {
                const P4<T> p = lds_p4(ts.tile + lds_u16(my_s + 2u * (uint32_t)r));
                const Key<T> key = Key<T>::make(dist2_rn<T, D>(ts.q.x, ts.q.y, ts.q.z, p.x, p.y, p.z), idx_of(p));
                if (r == K1 - 1) kth = key; else ok = ok && key.less(next);
                next = key;
                if (coop) sts_u32(my_s + 4u * (uint32_t)r, key.idx());
                else if (r >= drop) {
                    if (out32) out_idx32[row + r - drop] = key.idx() + 1u;
                    else out_idx[row + r - drop] = (int64_t)key.idx() + 1;
                    if (out_dist) out_dist[row + r - drop] = sqrt(key.d2());
                }
            }
            status = ts.accept(ok, kth);
        }
        if (coop) {
            __syncwarp();
            unsigned todo = __ballot_sync(FULL, status == TK_OK);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int64_t row_s = __shfl_sync(FULL, row, src);
                if (ts.lane < k_out) {
                    const uint32_t v = lds_u32(warp_s + (uint32_t)src * (uint32_t)(TK_LSTRIDE * 2) + 4u * (uint32_t)(ts.lane + drop)) + 1u;
                    if (out32) out_idx32[row_s + ts.lane] = v;
                    else out_idx[row_s + ts.lane] = (int64_t)v;
                }
            }
            __syncwarp();                                                                 // the lists are free for the next group
        }
        ts.report(status, fails);
    }
}

}  // namespace wtp
