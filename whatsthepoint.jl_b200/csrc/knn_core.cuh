// knn_core.cuh — the warp-per-query exact k-NN search on the uniform grid.
//
// One warp answers a run of consecutive sorted queries. For the cell of the current query
// the warp stages the 3^D block of neighbouring cells into its private shared-memory
// tile with TMA bulk copies (cp.async.bulk + mbarrier; the cells of one x-row are
// contiguous in the sorted array, so the block is <= 9 runs of 16/32-byte records, and
// every run is 16-byte aligned). Queries falling in the same cell reuse the tile. Each
// query sweeps the tile with one LDS.128 per candidate and keeps the K best in registers,
// distributed over the warp as a sorted list (rank = row*32 + lane) of canonical keys
// (d2, original index). f32 keys are packed into one 64-bit integer so a comparison is a
// single integer compare. After the sweep the K-th best is checked against the distance
// to the block's shell; only if it is not provably inside does the warp fall back to the
// general path, which walks further rings of cells straight from global memory with
// row/cell pruning. Blocks too large for the tile use the general path from ring 1.
//
// d2 = ((dx*dx + dy*dy) + dz*dz) in T with explicit round-to-nearest mul/add (never
// contracted into FMA), the arithmetic the CPU oracle and the reference's KD-tree use.
#pragma once
#include "common.cuh"

namespace wtp {

constexpr unsigned FULL = 0xffffffffu;

template <class T> __device__ __forceinline__ T t_inf();
template <> __device__ __forceinline__ float t_inf<float>() { return __int_as_float(0x7f800000); }
template <> __device__ __forceinline__ double t_inf<double>() { return __longlong_as_double(0x7ff0000000000000LL); }

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

template <class T, int D>
__device__ __forceinline__ T dist2_rn(T qx, T qy, T qz, T px, T py, T pz) {
    T dx = sub_rn(qx, px), dy = sub_rn(qy, py);
    T s = add_rn(mul_rn(dx, dx), mul_rn(dy, dy));
    if (D == 3) { T dz = sub_rn(qz, pz); s = add_rn(s, mul_rn(dz, dz)); }
    return s;
}

__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// The library's random unit vector for the pair (i, j) of one sweep (common.cuh: mix64, sweep_key): the stand-in for
// _safe_direction's randn(...)/norm (src/repel.jl:358-364). Cold path, kept out of line.
template <class T, int D>
__device__ __noinline__ void random_unit(uint64_t key, uint32_t i, uint32_t j, T* out /*[3]*/) {
    const uint64_t h1 = mix64(key ^ (((uint64_t)i << 32) | (uint64_t)j));
    for (uint64_t c = 0;; ++c) {
        const uint64_t h = mix64(h1 + c);
        const T scale = (T)(1.0 / 1048576.0);
        const T v0 = mul_rn((T)((int)(h & 0x1fffffu) - 1048576), scale);
        const T v1 = mul_rn((T)((int)((h >> 21) & 0x1fffffu) - 1048576), scale);
        const T v2 = D == 3 ? mul_rn((T)((int)((h >> 42) & 0x1fffffu) - 1048576), scale) : (T)0;
        T n2 = add_rn(mul_rn(v0, v0), mul_rn(v1, v1));
        if (D == 3) n2 = add_rn(n2, mul_rn(v2, v2));
        if (n2 <= (T)1 && n2 >= (T)(1.0 / 1024.0)) {
            const T n = sqrt_rn(n2);
            out[0] = div_rn(v0, n); out[1] = div_rn(v1, n); out[2] = D == 3 ? div_rn(v2, n) : (T)0;
            return;
        }
    }
}

template <class T>
__device__ __forceinline__ P4<T> load_p4(const P4<T>* p);
template <>
__device__ __forceinline__ P4<float> load_p4<float>(const P4<float>* p) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    P4<float> r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
}
template <>
__device__ __forceinline__ P4<double> load_p4<double>(const P4<double>* p) {
    double2 a = __ldg(reinterpret_cast<const double2*>(p));
    double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    P4<double> r; r.x = a.x; r.y = a.y; r.z = b.x; r.w = b.y; return r;
}
// shared-memory tile reads (LDS.128)
__device__ __forceinline__ P4<float> lds_p4(const P4<float>* p) {
    float4 v = *reinterpret_cast<const float4*>(p);
    P4<float> r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
}
__device__ __forceinline__ P4<double> lds_p4(const P4<double>* p) {
    double2 a = *reinterpret_cast<const double2*>(p);
    double2 b = *(reinterpret_cast<const double2*>(p) + 1);
    P4<double> r; r.x = a.x; r.y = a.y; r.z = b.x; r.w = b.y; return r;
}

// --------------------------------------------------------------- canonical keys
// Key<T>: total order (d2, idx). f32: one u64 = (bits(d2) << 32) | idx (d2 >= 0, so the
// IEEE bit pattern is monotone). f64: (double, u32) pair.
template <class T> struct Key;
template <> struct Key<float> {
    unsigned long long v;
    static __device__ __forceinline__ Key make(float d2, uint32_t idx) { Key k; k.v = ((unsigned long long)__float_as_uint(d2) << 32) | idx; return k; }
    static __device__ __forceinline__ Key sentinel() { Key k; k.v = 0x7f800000ffffffffULL; return k; }   // (+inf, 0xffffffff)
    __device__ __forceinline__ float d2() const { return __uint_as_float((uint32_t)(v >> 32)); }
    __device__ __forceinline__ uint32_t idx() const { return (uint32_t)v; }
    __device__ __forceinline__ bool less(const Key& o) const { return v < o.v; }
    __device__ __forceinline__ uint32_t coarse() const { return (uint32_t)(v >> 32); }    // monotone 32-bit image of d2
    __device__ __forceinline__ Key shfl(int src) const { Key k; k.v = __shfl_sync(FULL, v, src); return k; }
    __device__ __forceinline__ Key shfl_up1() const { Key k; k.v = __shfl_up_sync(FULL, v, 1); return k; }
    __device__ __forceinline__ Key shfl_xor(int m) const { Key k; k.v = __shfl_xor_sync(FULL, v, m); return k; }
};
template <> struct Key<double> {
    double d; uint32_t i;
    static __device__ __forceinline__ Key make(double d2, uint32_t idx) { Key k; k.d = d2; k.i = idx; return k; }
    static __device__ __forceinline__ Key sentinel() { Key k; k.d = t_inf<double>(); k.i = 0xffffffffu; return k; }
    __device__ __forceinline__ double d2() const { return d; }
    __device__ __forceinline__ uint32_t idx() const { return i; }
    __device__ __forceinline__ bool less(const Key& o) const { return d < o.d || (d == o.d && i < o.i); }
    __device__ __forceinline__ uint32_t coarse() const { return (uint32_t)__double2hiint(d); }   // d >= 0: high word is monotone
    __device__ __forceinline__ Key shfl(int src) const { Key k; k.d = __shfl_sync(FULL, d, src); k.i = __shfl_sync(FULL, i, src); return k; }
    __device__ __forceinline__ Key shfl_up1() const { Key k; k.d = __shfl_up_sync(FULL, d, 1); k.i = __shfl_up_sync(FULL, i, 1); return k; }
    __device__ __forceinline__ Key shfl_xor(int m) const { Key k; k.d = __shfl_xor_sync(FULL, d, m); k.i = __shfl_xor_sync(FULL, i, m); return k; }
};

// ------------------------------------------------------------ warp top-K list
template <class T, int KPL>
struct WarpList {
    Key<T> e[KPL];   // rank = row*32 + lane, ascending
    Key<T> thr;      // key at rank K-1 (warp-uniform)
    int K;

    __device__ __forceinline__ void init(int K_) {
        K = K_;
#pragma unroll
        for (int r = 0; r < KPL; ++r) e[r] = Key<T>::sentinel();
        thr = Key<T>::sentinel();
    }
    __device__ __forceinline__ void refresh_threshold() {
        const int e_thr = (K - 1) >> 5, l_thr = (K - 1) & 31;
        Key<T> t = e[0];
#pragma unroll
        for (int r = 1; r < KPL; ++r) if (r == e_thr) t = e[r];
        thr = t.shfl(l_thr);
    }
    // insert one warp-uniform candidate into the sorted list (the old last entry drops off)
    __device__ __forceinline__ void insert(const Key<T>& c, int lane) {
        Key<T> carry = c;
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
            Key<T> up = e[r].shfl_up1();
            Key<T> last = c;
            if (r + 1 < KPL) last = e[r].shfl(31);
            if (lane == 0) up = carry;
            const bool lt = c.less(e[r]);
            const bool lt_prev = (r == 0 && lane == 0) ? false : c.less(up);
            if (lt) e[r] = lt_prev ? up : c;
            carry = last;
        }
    }
    // offer one candidate per lane (sentinel on idle lanes)
    __device__ __forceinline__ void offer(const Key<T>& c, int lane) {
        unsigned m = __ballot_sync(FULL, c.less(thr));
        while (m) {
            const int b = __ffs(m) - 1;
            insert(c.shfl(b), lane);
            refresh_threshold();
            m = __ballot_sync(FULL, c.less(thr)) & ~((2u << b) - 1u);
        }
    }
    // 32 candidates (one per lane) into row 0 at once: bitonic-sort the batch, keep the 32
    // smallest of (row 0, batch) with one reversed elementwise min, then a 5-stage bitonic
    // merge. Cheaper than serial inserts once more than ~6 candidates qualify. KPL == 1 only.
    __device__ __forceinline__ void merge_sorted32(const Key<T>& c, int lane) {   // c: ascending over the lanes
        const Key<T> rb = c.shfl(31 - lane);
        if (rb.less(e[0])) e[0] = rb;
#pragma unroll
        for (int j = 16; j > 0; j >>= 1) {
            const Key<T> o = e[0].shfl_xor(j);
            const bool keep_min = (lane & j) == 0;
            if (o.less(e[0]) == keep_min) e[0] = o;
        }
    }
    static __device__ __forceinline__ Key<T> sort32(Key<T> c, int lane) {
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                const Key<T> o = c.shfl_xor(j);
                const bool keep_min = ((lane & j) == 0) == ((lane & k) == 0);
                if (o.less(c) == keep_min) c = o;   // equal keys: either choice is the same value
            }
        }
        return c;
    }
    // Same result as sort32, cheaper: sort 32-bit images (the top 27 bits of d2's pattern, the lane in
    // the low 5) with one SHFL + one min/max per stage, fetch the full keys through the sorted lane
    // numbers, and check that neighbours are in strict canonical order. The images of two keys collide
    // only when their d2 agree to ~18 bits (or on exact ties); then the order of that pair is decided
    // by the full sort instead.
    static __device__ __forceinline__ Key<T> sort32_fast(const Key<T>& c, int lane) {
        uint32_t k = (c.coarse() & ~31u) | (uint32_t)lane;
#pragma unroll
        for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
            for (int j = kk >> 1; j > 0; j >>= 1) {
                const uint32_t o = __shfl_xor_sync(FULL, k, j);
                const bool keep_min = ((lane & j) == 0) == ((lane & kk) == 0);
                k = keep_min ? min(k, o) : max(k, o);
            }
        }
        const Key<T> s = c.shfl((int)(k & 31u));
        const Key<T> prev = s.shfl_up1();
        const bool bad = lane > 0 && !prev.less(s) && !(s.idx() == 0xffffffffu && prev.idx() == 0xffffffffu);   // sentinels tie with each other
        if (__ballot_sync(FULL, bad)) return sort32(c, lane);
        return s;
    }
    // first batch into an empty list: bitonic sort of the 32 candidates straight into row 0
    __device__ __forceinline__ void seed(Key<T> c, int lane) {
        e[0] = sort32(c, lane);
        refresh_threshold();
    }
};

// ------------------------------------------------------------------ TMA helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// shared-memory accesses that must keep their program order among themselves (volatile asm): per-thread lists that
// are written and read back through differently typed views
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return (uint32_t)v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v)); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v)); }

// row offsets of the 3^D block, centre row first, then face rows, then corner rows
__device__ __forceinline__ int row_dy(int t) { return (int)((0x22161u >> (2 * t)) & 3u) - 1; }   // 0,-1,+1,0,0,-1,+1,-1,+1
__device__ __forceinline__ int row_dz(int t) { return (int)((0x28215u >> (2 * t)) & 3u) - 1; }   // 0,0,0,-1,+1,-1,-1,+1,+1

// ------------------------------------------------------------- the search
// TILE_CAP: points per warp tile (0 disables staging: general path only).
template <class T, int D, int KPL, int TILE_CAP>
struct WarpKnn {
    static constexpr int NROWS = D == 3 ? 9 : 3;
    const Grid<T>& g;
    const P4<T>* __restrict__ sorted;
    const uint32_t* __restrict__ cell_start;
    P4<T>* tile;        // this warp's shared-memory tile (TILE_CAP records)
    Key<T>* buf;        // this warp's pre-filter buffer (PF_CAP keys)
    T r0sq;             // pre-filter radius^2 for the staged block (0: disabled)
    uint64_t* bar;      // this warp's mbarrier
    uint32_t phase;
    int lane;
    int scx, scy, scz;  // staged cell (-1: none)
    uint32_t tile_n;    // staged candidates; 0xffffffff: block does not fit the tile
    uint32_t block_n;   // points in the 3^D block of the staged cell
    uint32_t row_begin, row_len;   // lane t < NROWS: its row of the block in the sorted array
    T qx, qy, qz;
    int cx, cy, cz;
    bool seeded;
    bool missed;        // a ring reached a layer outside the index window (Grid::w_lo, w_hi): the list may be incomplete
    WarpList<T, KPL> list;

    static constexpr int PF_CAP = 64;
    __device__ __forceinline__ WarpKnn(const Grid<T>& g_, const P4<T>* s, const uint32_t* cs, P4<T>* tile_, Key<T>* buf_, uint64_t* bar_, int lane_)
        : g(g_), sorted(s), cell_start(cs), tile(tile_), buf(buf_), r0sq((T)0), bar(bar_), phase(0), lane(lane_), scx(-1), scy(-1), scz(-1), tile_n(0), block_n(0), row_begin(0), row_len(0) {
        if (TILE_CAP > 0) {
            if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
            __syncwarp();
        }
    }

    __device__ __forceinline__ T face(int d, int j) const { return add_rn(g.lo[d], mul_rn((T)j, g.c)); }
    // conservative gap from the query to the slab of cells at offset o (o != 0) along d
    __device__ __forceinline__ T gap(int d, T q, int c0, int o) const {
        if (o == 0) return (T)0;
        T gp = o < 0 ? sub_rn(q, face(d, c0 + o + 1)) : sub_rn(face(d, c0 + o), q);
        gp = sub_rn(gp, g.slack);
        return gp > (T)0 ? gp : (T)0;
    }
    __device__ __forceinline__ T lb3(T gx, T gy, T gz) const {
        T s = add_rn(mul_rn(gx, gx), mul_rn(gy, gy));
        if (D == 3) s = add_rn(s, mul_rn(gz, gz));
        return s;
    }
    __device__ __forceinline__ void consume(bool valid, const P4<T>& p) {
        Key<T> c = Key<T>::sentinel();
        if (valid) c = Key<T>::make(dist2_rn<T, D>(qx, qy, qz, p.x, p.y, p.z), idx_of(p));
        if (!seeded) { list.seed(c, lane); seeded = true; }
        else list.offer(c, lane);
    }

    // ---- staged path ---------------------------------------------------------
    // Lanes 0..NROWS-1 each resolve one x-row of the block (two cell_start loads in
    // parallel), a warp scan lays the rows out back to back in the tile, and each of those
    // lanes issues the bulk copy of its own row; everyone waits on the warp's mbarrier.
    __device__ __forceinline__ void stage() {
        uint32_t begin = 0, len = 0;
        if (lane < NROWS) {
            const int ry = cy + row_dy(lane), rz = D == 3 ? cz + row_dz(lane) : 0;
            if (ry >= 0 && ry < g.n[1] && rz >= 0 && rz < g.n[2]) {
                const int x0 = cx > 0 ? cx - 1 : 0, x1 = cx < g.n[0] - 1 ? cx + 1 : g.n[0] - 1;
                const uint32_t row = ((uint32_t)rz * (uint32_t)g.n[1] + (uint32_t)ry) * (uint32_t)g.n[0];
                begin = cell_start[row + x0];
                len = cell_start[row + x1 + 1] - begin;
            }
        }
        uint32_t incl = len;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t total = __shfl_sync(FULL, incl, NROWS - 1);
        scx = cx; scy = cy; scz = cz;
        block_n = total; row_begin = begin; row_len = len;
        if (total > (uint32_t)TILE_CAP) { tile_n = 0xffffffffu; return; }
        tile_n = total;
        if (total == 0) return;
        __syncwarp();   // every lane is done reading the previous tile
        if (lane == 0) mbar_expect_tx(bar, total * (uint32_t)sizeof(P4<T>));
        __syncwarp();
        if (len > 0) tma_bulk_g2s(tile + (incl - len), sorted + begin, len * (uint32_t)sizeof(P4<T>), bar);
        mbar_wait(bar, phase);
        phase ^= 1u;
    }
    // Pre-filter radius for the staged block: the local density is tile_n points in the
    // block, so a ball holding `target` points has r^D = target * V_block / (V_unit_ball * tile_n).
    // Only a guess — exactness never depends on it (prefilter_tile reports failure).
    __device__ __forceinline__ void set_prefilter_radius(int K) {
        const float target = (float)K + 2.5f * sqrtf((float)K) + 1.0f;
        const float c = (float)g.c;
        float r2;
        if (D == 3) { const float r3 = target * 27.0f / (4.18879f * (float)block_n); r2 = c * c * cbrtf(r3 * r3); }
        else r2 = c * c * target * 9.0f / (3.14159265f * (float)block_n);
        r0sq = (T)r2;
    }
    // Pre-filter: every candidate with d2 <= r0sq goes to the warp's buffer. If between K and
    // PF_CAP candidates qualify, the K best of the block are among them: one bitonic sort
    // (plus a merge of the second half) builds the whole list.
    __device__ __forceinline__ void prefilter_push(bool valid, const P4<T>& p, uint32_t& cnt) {
        bool q = false;
        Key<T> c = Key<T>::sentinel();
        if (valid) {
            const T d = dist2_rn<T, D>(qx, qy, qz, p.x, p.y, p.z);
            c = Key<T>::make(d, idx_of(p));
            q = !(d > r0sq);
        }
        const unsigned m = __ballot_sync(FULL, q);
        const uint32_t pos = cnt + __popc(m & ((1u << lane) - 1u));
        if (q && pos < (uint32_t)PF_CAP) buf[pos] = c;
        cnt += __popc(m);
    }
    __device__ __forceinline__ bool prefilter_finish(int K, uint32_t cnt) {
        if (cnt < (uint32_t)K || cnt > (uint32_t)PF_CAP) return false;
        __syncwarp();
        // First half: bitonic sort straight into the list. Second half (if any): a handful of
        // extras are inserted one by one, more are sorted (same code copy, second trip) and merged.
#pragma unroll 1
        for (uint32_t h = 0; h < cnt; h += 32) {
            Key<T> c = h + (uint32_t)lane < cnt ? buf[h + lane] : Key<T>::sentinel();
            if (h == 0 || cnt > 38u) {
                c = WarpList<T, KPL>::sort32_fast(c, lane);
                if (h == 0) list.e[0] = c;
                else list.merge_sorted32(c, lane);
            } else {
                list.refresh_threshold();
                list.offer(c, lane);
            }
        }
        list.refresh_threshold();
        __syncwarp();
        seeded = true;
        return true;
    }
    __device__ __forceinline__ bool prefilter_tile(int K) {
        uint32_t cnt = 0;
        for (uint32_t j0 = 0; j0 < tile_n; j0 += 32) {
            const uint32_t j = j0 + lane;
            const bool valid = j < tile_n;
            P4<T> p;
            if (valid) p = lds_p4(tile + j);
            prefilter_push(valid, p, cnt);
        }
        return prefilter_finish(K, cnt);
    }
    // Same filter for blocks too large for the tile (dense regions of graded clouds): rows are
    // read straight from global memory, and rows / end cells farther than the filter radius
    // are skipped.
    __device__ __forceinline__ bool prefilter_rows(int K) {
        uint32_t cnt = 0;
#pragma unroll 1
        for (int t = 0; t < NROWS; ++t) {
            const uint32_t begin = __shfl_sync(FULL, row_begin, t), len = __shfl_sync(FULL, row_len, t);
            if (len == 0) continue;
            const T gy = gap(1, qy, cy, row_dy(t)), gz = D == 3 ? gap(2, qz, cz, row_dz(t)) : (T)0;
            if (lb3((T)0, gy, gz) > r0sq) continue;
            for (uint32_t j0 = begin; j0 < begin + len && cnt <= (uint32_t)PF_CAP; j0 += 32) {
                const uint32_t j = j0 + lane;
                const bool valid = j < begin + len;
                P4<T> p;
                if (valid) p = load_p4<T>(sorted + j);
                prefilter_push(valid, p, cnt);
            }
        }
        return prefilter_finish(K, cnt);
    }
    // classic sweep of the staged tile (pre-filter count out of range): seed batch, then inserts
    __device__ __forceinline__ void sweep_tile() {
        for (uint32_t j0 = 0; j0 < tile_n; j0 += 32) {
            const uint32_t j = j0 + lane;
            const bool valid = j < tile_n;
            P4<T> p;
            if (valid) p = lds_p4(tile + j);
            consume(valid, p);
        }
    }

    // ---- general path --------------------------------------------------------
    __device__ __forceinline__ void sweep(uint32_t begin, uint32_t end) {
        for (uint32_t j0 = begin; j0 < end; j0 += 32) {
            const uint32_t j = j0 + lane;
            const bool valid = j < end;
            P4<T> p;
            if (valid) p = load_p4<T>(sorted + j);
            consume(valid, p);
        }
    }
    __device__ __forceinline__ void sweep_cells(int x0, int x1, int ry, int rz) {
        const uint32_t row = ((uint32_t)rz * (uint32_t)g.n[1] + (uint32_t)ry) * (uint32_t)g.n[0];
        sweep(cell_start[row + x0], cell_start[row + x1 + 1]);
    }
    // Row at offset (dy, dz) of ring R: the whole x-span if the row is on the ring's outer
    // shell (or R == 1), else just the two end cells; pruned by distance lower bounds.
    __device__ __forceinline__ void ring_row(int R, int dy, int dz) {
        const int ry = cy + dy, rz = cz + dz;
        if (ry < 0 || ry >= g.n[1] || rz < 0 || rz >= g.n[2]) return;
        const int layer = D == 3 ? rz : ry;
        if (layer < g.w_lo || layer > g.w_hi) { missed = true; return; }   // not indexed here (windowed index): reported, never read
        const T gy = gap(1, qy, cy, dy), gz = D == 3 ? gap(2, qz, cz, dz) : (T)0;
        const int ady = dy < 0 ? -dy : dy, adz = dz < 0 ? -dz : dz;
        const bool outer = (ady > adz ? ady : adz) == R;
        const T thr_d = list.thr.d2();
        if (outer || R == 1) {
            if (lb3((T)0, gy, gz) > thr_d) return;
            int x0 = cx - R, x1 = cx + R;
            if (x0 < 0) x0 = 0; else if (lb3(gap(0, qx, cx, -R), gy, gz) > thr_d) ++x0;
            if (x1 > g.n[0] - 1) x1 = g.n[0] - 1; else if (lb3(gap(0, qx, cx, R), gy, gz) > thr_d) --x1;
            if (x0 <= x1) sweep_cells(x0, x1, ry, rz);
        } else {
            if (cx - R >= 0 && !(lb3(gap(0, qx, cx, -R), gy, gz) > thr_d)) sweep_cells(cx - R, cx - R, ry, rz);
            if (cx + R <= g.n[0] - 1 && !(lb3(gap(0, qx, cx, R), gy, gz) > list.thr.d2())) sweep_cells(cx + R, cx + R, ry, rz);
        }
    }
    // true when every point outside the block of radius R is provably worse than the K-th best
    __device__ __forceinline__ bool block_is_exact(int R) const {
        T shell = t_inf<T>();
        bool open = false;
        if (cx - R > 0) { open = true; shell = fmin(shell, sub_rn(sub_rn(qx, face(0, cx - R)), g.slack)); }
        if (cx + R < g.n[0] - 1) { open = true; shell = fmin(shell, sub_rn(sub_rn(face(0, cx + R + 1), qx), g.slack)); }
        if (cy - R > 0) { open = true; shell = fmin(shell, sub_rn(sub_rn(qy, face(1, cy - R)), g.slack)); }
        if (cy + R < g.n[1] - 1) { open = true; shell = fmin(shell, sub_rn(sub_rn(face(1, cy + R + 1), qy), g.slack)); }
        if (D == 3) {
            if (cz - R > 0) { open = true; shell = fmin(shell, sub_rn(sub_rn(qz, face(2, cz - R)), g.slack)); }
            if (cz + R < g.n[2] - 1) { open = true; shell = fmin(shell, sub_rn(sub_rn(face(2, cz + R + 1), qz), g.slack)); }
        }
        if (!open) return true;            // the block covers the whole grid
        if (!(shell > (T)0)) return false;
        return list.thr.d2() < mul_rn(shell, shell);
    }

    // Runs the search for one query; returns the number of rings swept (1 = the 3^D block sufficed).
    __device__ __forceinline__ int run(T x, T y, T z, int K) {
        qx = x; qy = y; qz = z;
        cx = cell_coord(g, qx, 0);
        cy = cell_coord(g, qy, 1);
        cz = D == 3 ? cell_coord(g, qz, 2) : 0;
        seeded = false;
        missed = false;
        list.init(K);
        bool staged = false;
        if (TILE_CAP > 0) {
            if (cx != scx || cy != scy || cz != scz) stage();
            staged = tile_n != 0xffffffffu;
        }
        bool done = false;
        if (TILE_CAP > 0 && KPL == 1 && block_n >= (uint32_t)K) {
            set_prefilter_radius(K);
            done = staged ? prefilter_tile(K) : prefilter_rows(K);
        }
        // General path (one code copy): rings of cells from global memory with pruning. Ring 1 is
        // swept here only when the pre-filter did not settle the 3^D block.
        if (!done && staged) { sweep_tile(); done = true; }
        int R = done ? 1 : 0;
#pragma unroll 1
        while (R == 0 || !block_is_exact(R)) {
            ++R;
            const int Rz = D == 3 ? R : 0;
#pragma unroll 1
            for (int dz = -Rz; dz <= Rz; ++dz)
#pragma unroll 1
                for (int dy = -R; dy <= R; ++dy) ring_row(R, dy, dz);
        }
        return R;
    }
};

}  // namespace wtp
