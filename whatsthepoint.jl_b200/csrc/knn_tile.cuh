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
constexpr int TK_LSTRIDE = 50;     // u16 slots per list: 25 words, odd, so the 32 lists of a warp start in 32 different banks
template <class T> __host__ __device__ constexpr int tk_cap() { return sizeof(T) == 4 ? 1792 : 1664; }   // records per CTA tile
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
    const float target = (float)K + 2.5f * sqrtf((float)K) + 1.0f;
    const float c = (float)g.c;
    float r2;
    if (D == 3) { const float r3 = target * 27.0f / (4.18879f * (float)block_n); r2 = c * c * cbrtf(r3 * r3); }
    else r2 = c * c * target * 9.0f / (3.14159265f * (float)block_n);
    return (T)r2;
}

template <class T, int D>
__global__ void __launch_bounds__(TK_Q, sizeof(T) == 4 ? 5 : 3)
knn_tile_kernel(const Grid<T> g, const P4<T>* __restrict__ sorted, const uint32_t* __restrict__ cell_start, uint32_t s_begin, uint32_t s_end,
                uint32_t q_begin, int K1, int drop, void* __restrict__ out_idx_v, int out32, T* __restrict__ out_dist,
                uint32_t* __restrict__ fail_list, uint32_t* __restrict__ fail_count) {
    constexpr int NROWS = D == 3 ? 9 : 3;
    constexpr int CAP = tk_cap<T>();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    P4<T>* tile = reinterpret_cast<P4<T>*>(smem_raw);
    uint16_t* lists = reinterpret_cast<uint16_t*>(smem_raw + (size_t)CAP * sizeof(P4<T>));   // [thread][slot]
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_rowid[TK_Q];
    __shared__ uint32_t s_next, s_total;
    __shared__ int s_x0, s_x1;
    __shared__ uint32_t s_run_begin[NROWS], s_run_off[NROWS], s_run_base[NROWS];   // base: first cell id of the row (0xffffffff: outside)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t j = s_begin + blockIdx.x * TK_Q + tid;
    const bool active = j < s_end;
    P4<T> q; q.x = q.y = q.z = (T)0; q.w = idx_bits((T)0, 0u);
    int cx = 0, cy = 0, cz = 0;
    if (active) {
        q = load_p4<T>(sorted + j);
        cx = cell_coord(g, q.x, 0); cy = cell_coord(g, q.y, 1); cz = D == 3 ? cell_coord(g, q.z, 2) : 0;
    }
    const uint32_t rowid = active ? (uint32_t)cz * (uint32_t)g.n[1] + (uint32_t)cy : 0xffffffffu;
    s_rowid[tid] = rowid;
    if (tid == 0) { s_next = 0; mbar_init(&s_bar, 1); fence_mbar_init(); }
    const T shell2 = block_shell2<T, D>(g, q.x, q.y, q.z, cx, cy, cz);
    const int k_out = K1 - drop;
    int64_t* __restrict__ out_idx = static_cast<int64_t*>(out_idx_v);
    uint32_t* __restrict__ out_idx32 = static_cast<uint32_t*>(out_idx_v);
    uint32_t phase = 0;

    for (;;) {
        __syncthreads();                                   // previous group is done with the tile; s_next is final
        const uint32_t first = s_next;
        if (first >= TK_Q || s_rowid[first] == 0xffffffffu) break;
        const uint32_t grp_row = s_rowid[first];
        const bool in_group = active && (uint32_t)tid >= first && rowid == grp_row;
        if ((uint32_t)tid == first) s_x0 = cx;
        if (in_group && (tid == TK_Q - 1 || s_rowid[tid + 1] != grp_row)) { s_x1 = cx; }
        __syncthreads();
        // ---- stage the slab: rows (cy+dy, cz+dz), cells [x0-1, x1+1]
        if (warp == 0) {
            uint32_t begin = 0, len = 0, base = 0xffffffffu;
            if (lane < NROWS) {
                const int gy = (int)(grp_row % (uint32_t)g.n[1]), gz = (int)(grp_row / (uint32_t)g.n[1]);
                const int ry = gy + row_dy(lane), rz = D == 3 ? gz + row_dz(lane) : 0;
                if (ry >= 0 && ry < g.n[1] && rz >= 0 && rz < g.n[2]) {
                    const int x0 = s_x0 > 0 ? s_x0 - 1 : 0, x1 = s_x1 < g.n[0] - 1 ? s_x1 + 1 : g.n[0] - 1;
                    base = ((uint32_t)rz * (uint32_t)g.n[1] + (uint32_t)ry) * (uint32_t)g.n[0];
                    begin = cell_start[base + x0];
                    len = cell_start[base + x1 + 1] - begin;
                }
            }
            uint32_t incl = len;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t total = __shfl_sync(FULL, incl, NROWS - 1);
            if (lane < NROWS) { s_run_begin[lane] = begin; s_run_off[lane] = incl - len; s_run_base[lane] = base; }
            if (lane == 0) { s_total = total; if (total > 0 && total <= (uint32_t)CAP) mbar_expect_tx(&s_bar, total * (uint32_t)sizeof(P4<T>)); }
            __syncwarp();
            if (total > 0 && total <= (uint32_t)CAP && len > 0) tma_bulk_g2s(tile + (incl - len), sorted + begin, len * (uint32_t)sizeof(P4<T>), &s_bar);
        }
        // the last member of the group hands over to the next group
        if (in_group && (tid == TK_Q - 1 || s_rowid[tid + 1] != grp_row)) s_next = (uint32_t)tid + 1;
        __syncthreads();
        const uint32_t total = s_total;
        const bool fits = total <= (uint32_t)CAP;
        if (fits && total > 0) { mbar_wait(&s_bar, phase); phase ^= 1u; }

        // ---- phase A: thread per query. The filter uses a fused (cheaper) distance and a radius padded by a
        // few ulps, so the list holds at least every block point with canonical d2 <= r0sq.
        uint32_t cnt = 0;
        bool fail = !fits;
        T r0sq = (T)0;
        uint16_t* my = lists + tid * TK_LSTRIDE;
        if (in_group && fits) {
            uint32_t b[NROWS], e[NROWS], block_n = 0;
            const int xa = cx > 0 ? cx - 1 : 0, xb = cx < g.n[0] - 1 ? cx + 1 : g.n[0] - 1;
#pragma unroll
            for (int r = 0; r < NROWS; ++r) {
                const uint32_t base = s_run_base[r];
                b[r] = e[r] = 0;
                if (base != 0xffffffffu) {
                    const uint32_t shift = s_run_off[r] - s_run_begin[r];
                    b[r] = __ldg(cell_start + base + xa) + shift;
                    e[r] = __ldg(cell_start + base + xb + 1) + shift;
                    block_n += e[r] - b[r];
                }
            }
            if (block_n < (uint32_t)K1) fail = true;
            else {
                r0sq = prefilter_radius2<T, D>(g, block_n, K1);
                const T r0pad = r0sq * ((T)1 + (T)8 * (sizeof(T) == 4 ? (T)1.1920929e-7 : (T)2.220446049250313e-16));
                const uint32_t my_s = smem_u32(my);
#pragma unroll
                for (int r = 0; r < NROWS; ++r) {
#pragma unroll 2
                    for (uint32_t t = b[r]; t < e[r]; ++t) {
                        const P4<T> p = lds_p4(tile + t);
                        const T dx = q.x - p.x, dy = q.y - p.y;
                        T d = fma(dy, dy, dx * dx);
                        if (D == 3) { const T dz = q.z - p.z; d = fma(dz, dz, d); }
                        const uint32_t slot = cnt < (uint32_t)TK_LCAP ? cnt : (uint32_t)TK_LCAP;
                        const uint32_t hit = d <= r0pad ? 1u : 0u;
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u16 [%0], %1;\n\t}"
                                     ::"r"(my_s + slot * 2u), "h"((uint16_t)t), "r"(hit) : "memory");
                        cnt += hit;
                    }
                }
                if (cnt < (uint32_t)K1 || cnt > (uint32_t)TK_LCAP) fail = true;
            }
        }
        // ---- phase B: still thread per query. Canonical keys of the hits, reduced to 32-bit images (top 26
        // bits of d2's pattern, the list slot below), go through a fixed sorting network in registers; the
        // K smallest are then re-read in that order, checked for strict canonical order (the images of two
        // keys collide only when their d2 agree to ~17 bits, or on exact ties) and written out.
        if (in_group && !fail) {
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
            const int64_t row = (int64_t)(idx_of(q) - q_begin) * k_out;
            Key<T> prev = Key<T>::make((T)0, 0u);
            bool ok = true;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                if (r < K1) {
                    const P4<T> p = lds_p4(tile + my[k[r] & 63u]);
                    const Key<T> key = Key<T>::make(dist2_rn<T, D>(q.x, q.y, q.z, p.x, p.y, p.z), idx_of(p));
                    if (r > 0) ok = ok && prev.less(key);
                    prev = key;
                    if (r >= drop) {
                        if (out32) out_idx32[row + r - drop] = key.idx() + 1u;
                        else out_idx[row + r - drop] = (int64_t)key.idx() + 1;
                        if (out_dist) out_dist[row + r - drop] = sqrt(key.d2());
                    }
                }
                if (r + 1 == K1) ok = ok && (k[r] >> 6) != (k[r + 1] >> 6);   // the first key left out must not be a look-alike
            }
            // prev is the K-th key: it must lie inside the guaranteed radius of the list and inside the block's shell
            fail = !(ok && prev.d2() <= r0sq && prev.d2() < shell2);
        }
        const unsigned failed = __ballot_sync(FULL, in_group && fail);
        if (failed) {                                        // hand the rest to the general kernel
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(fail_count, (uint32_t)__popc(failed));
            base = __shfl_sync(FULL, base, 0);
            if ((failed >> lane) & 1u) fail_list[base + __popc(failed & ((1u << lane) - 1u))] = j;
        }
    }
}

}  // namespace wtp
