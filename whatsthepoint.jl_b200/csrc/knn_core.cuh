// knn_core.cuh — the warp-per-query exact k-NN search on the uniform grid.
//
// One warp answers one query. The K best candidates live in registers, distributed
// over the warp as a sorted list (rank = row*32 + lane, KPL rows per lane) ordered by
// the canonical key (d2, original index). The warp sweeps the 3^D block of cells around
// the query row by row (cells of one x-row are contiguous in the sorted array, so a row
// is one coalesced run of 16/32-byte records), prunes rows and end cells whose distance
// lower bound already exceeds the current K-th best, and keeps expanding ring by ring
// until the K-th best is provably inside the swept block.
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

template <class T>
__device__ __forceinline__ bool key_less(T ad, uint32_t ai, T bd, uint32_t bi) {
    return ad < bd || (ad == bd && ai < bi);
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

// ------------------------------------------------------------ warp top-K list
template <class T, int KPL>
struct WarpList {
    T d2[KPL];
    uint32_t idx[KPL];
    T thr_d;          // key at rank K-1 (warp-uniform)
    uint32_t thr_i;
    int K;

    __device__ __forceinline__ void init(int K_) {
        K = K_;
#pragma unroll
        for (int e = 0; e < KPL; ++e) { d2[e] = t_inf<T>(); idx[e] = 0xffffffffu; }
        thr_d = t_inf<T>(); thr_i = 0xffffffffu;
    }
    __device__ __forceinline__ void refresh_threshold() {
        const int e_thr = (K - 1) >> 5, l_thr = (K - 1) & 31;
        T td = d2[0]; uint32_t ti = idx[0];
#pragma unroll
        for (int e = 1; e < KPL; ++e) if (e == e_thr) { td = d2[e]; ti = idx[e]; }
        thr_d = __shfl_sync(FULL, td, l_thr);
        thr_i = __shfl_sync(FULL, ti, l_thr);
    }
    // insert one warp-uniform candidate into the sorted list (drops the old last entry)
    __device__ __forceinline__ void insert(T cd, uint32_t ci, int lane) {
        T carry_d = (T)0; uint32_t carry_i = 0;
#pragma unroll
        for (int e = 0; e < KPL; ++e) {
            T up_d = __shfl_up_sync(FULL, d2[e], 1);
            uint32_t up_i = __shfl_up_sync(FULL, idx[e], 1);
            T last_d = (T)0; uint32_t last_i = 0;
            if (e + 1 < KPL) { last_d = __shfl_sync(FULL, d2[e], 31); last_i = __shfl_sync(FULL, idx[e], 31); }
            if (lane == 0) { up_d = carry_d; up_i = carry_i; }
            const bool lt = key_less(cd, ci, d2[e], idx[e]);
            const bool lt_prev = (e == 0 && lane == 0) ? false : key_less(cd, ci, up_d, up_i);
            if (lt) { d2[e] = lt_prev ? up_d : cd; idx[e] = lt_prev ? up_i : ci; }
            carry_d = last_d; carry_i = last_i;
        }
    }
    // offer one candidate per lane (cd = +inf on idle lanes)
    __device__ __forceinline__ void offer(T cd, uint32_t ci, int lane) {
        unsigned m = __ballot_sync(FULL, key_less(cd, ci, thr_d, thr_i));
        while (m) {
            const int b = __ffs(m) - 1;
            const T bd = __shfl_sync(FULL, cd, b);
            const uint32_t bi = __shfl_sync(FULL, ci, b);
            insert(bd, bi, lane);
            refresh_threshold();
            m = __ballot_sync(FULL, key_less(cd, ci, thr_d, thr_i)) & ~((2u << b) - 1u);
        }
    }
    // first batch into an empty list: bitonic sort of the 32 candidates straight into row 0
    __device__ __forceinline__ void seed(T cd, uint32_t ci, int lane) {
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                const T od = __shfl_xor_sync(FULL, cd, j);
                const uint32_t oi = __shfl_xor_sync(FULL, ci, j);
                const bool keep_min = ((lane & j) == 0) == ((lane & k) == 0);
                const bool other_less = key_less(od, oi, cd, ci);
                const bool mine_less = key_less(cd, ci, od, oi);
                const bool take = keep_min ? other_less : mine_less;
                if (take) { cd = od; ci = oi; }
            }
        }
        d2[0] = cd; idx[0] = ci;
        refresh_threshold();
    }
};

// ------------------------------------------------------------- the search
template <class T, int D, int KPL>
struct WarpKnn {
    const Grid<T>& g;
    const P4<T>* __restrict__ sorted;
    const uint32_t* __restrict__ cell_start;
    T qx, qy, qz;
    int cx, cy, cz, lane;
    bool seeded;
    WarpList<T, KPL> list;

    __device__ __forceinline__ WarpKnn(const Grid<T>& g_, const P4<T>* s, const uint32_t* cs) : g(g_), sorted(s), cell_start(cs) {}

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
    __device__ __forceinline__ void sweep(uint32_t begin, uint32_t end) {
        for (uint32_t j0 = begin; j0 < end; j0 += 32) {
            const uint32_t j = j0 + lane;
            T cd = t_inf<T>(); uint32_t ci = 0xffffffffu;
            if (j < end) {
                const P4<T> p = load_p4<T>(sorted + j);
                cd = dist2_rn<T, D>(qx, qy, qz, p.x, p.y, p.z);
                ci = idx_of(p);
            }
            if (!seeded) { list.seed(cd, ci, lane); seeded = true; }
            else list.offer(cd, ci, lane);
        }
    }
    // cells [x0, x1] of row (ry, rz); the caller has clamped and pruned
    __device__ __forceinline__ void sweep_cells(int x0, int x1, int ry, int rz) {
        const uint32_t row = ((uint32_t)rz * (uint32_t)g.n[1] + (uint32_t)ry) * (uint32_t)g.n[0];
        sweep(cell_start[row + x0], cell_start[row + x1 + 1]);
    }
    // Row at offset (dy, dz) of ring R: the whole x-span if the row is on the ring's outer
    // shell (or R == 1 centre row), else just the two end cells.
    __device__ __forceinline__ void ring_row(int R, int dy, int dz) {
        const int ry = cy + dy, rz = cz + dz;
        if (ry < 0 || ry >= g.n[1] || rz < 0 || rz >= g.n[2]) return;
        const T gy = gap(1, qy, cy, dy), gz = D == 3 ? gap(2, qz, cz, dz) : (T)0;
        const int ady = dy < 0 ? -dy : dy, adz = dz < 0 ? -dz : dz;
        const bool outer = (ady > adz ? ady : adz) == R;
        if (outer || R == 1) {
            if (lb3((T)0, gy, gz) > list.thr_d) return;
            int x0 = cx - R, x1 = cx + R;
            if (x0 < 0) x0 = 0; else if (lb3(gap(0, qx, cx, -R), gy, gz) > list.thr_d) ++x0;
            if (x1 > g.n[0] - 1) x1 = g.n[0] - 1; else if (lb3(gap(0, qx, cx, R), gy, gz) > list.thr_d) --x1;
            if (x0 <= x1) sweep_cells(x0, x1, ry, rz);
        } else {
            if (cx - R >= 0 && !(lb3(gap(0, qx, cx, -R), gy, gz) > list.thr_d)) sweep_cells(cx - R, cx - R, ry, rz);
            if (cx + R <= g.n[0] - 1 && !(lb3(gap(0, qx, cx, R), gy, gz) > list.thr_d)) sweep_cells(cx + R, cx + R, ry, rz);
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
        return list.thr_d < mul_rn(shell, shell);
    }
    // Runs the search; returns the number of rings swept (1 = the 3^D block sufficed).
    __device__ __forceinline__ int run(T x, T y, T z, int K, int lane_) {
        qx = x; qy = y; qz = z; lane = lane_;
        cx = cell_coord(g, qx, 0);
        cy = cell_coord(g, qy, 1);
        cz = D == 3 ? cell_coord(g, qz, 2) : 0;
        seeded = false;
        list.init(K);
        // ring 1, centre row first, then face rows, then corner rows
        ring_row(1, 0, 0);
        ring_row(1, -1, 0); ring_row(1, 1, 0);
        if (D == 3) {
            ring_row(1, 0, -1); ring_row(1, 0, 1);
            ring_row(1, -1, -1); ring_row(1, 1, -1); ring_row(1, -1, 1); ring_row(1, 1, 1);
        }
        int R = 1;
        while (!block_is_exact(R)) {
            ++R;
            const int Rz = D == 3 ? R : 0;
            for (int dz = -Rz; dz <= Rz; ++dz)
                for (int dy = -R; dy <= R; ++dy) ring_row(R, dy, dz);
        }
        return R;
    }
};

}  // namespace wtp
