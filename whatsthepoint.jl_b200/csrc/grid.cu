// grid.cu — the device spatial index: bounding box, cell keys, LSD radix sort with one
// byte (8-bit digit) per pass, exclusive scans, and the gather into sorted 16/32-byte
// point records with per-cell start offsets.
//
// Replaces the per-call KDTree construction of the reference (Meshes KNearestSearch /
// BallSearch at src/topology.jl:80,93 and KDTree(coords) at src/repel.jl:218,252).
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "kernels.cuh"

namespace wtp {

#define LAUNCH_CHECK(ctx)                         \
    do {                                          \
        (ctx)->launches++;                        \
        WTP_CUDA_CHECK(cudaPeekAtLastError());    \
    } while (0)

// =========================================================== bounding box
template <class T> struct Lim;
template <> struct Lim<float> { static __host__ __device__ float inf() { return __builtin_huge_valf(); } };
template <> struct Lim<double> { static __host__ __device__ double inf() { return __builtin_huge_val(); } };

constexpr int BBOX_THREADS = 256;

template <class T> struct Vec16;
template <> struct Vec16<float> { using type = float4; static constexpr int n = 4; };
template <> struct Vec16<double> { using type = double2; static constexpr int n = 2; };
__device__ inline void unpack16(const float4& v, float* e) { e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w; }
__device__ inline void unpack16(const double2& v, double* e) { e[0] = v.x; e[1] = v.y; }

// The coordinates are read as one flat stream of 16-byte vectors. The grid stride is a multiple of D elements
// (gridDim.x is a multiple of 3), so component c of a thread's vector belongs to the same axis in every iteration.
template <class T, int D>
__global__ void __launch_bounds__(BBOX_THREADS) bbox_partial_kernel(const T* __restrict__ pts, int64_t N, T* __restrict__ partial) {
    using V = typename Vec16<T>::type;
    constexpr int VN = Vec16<T>::n;
    const int64_t F = N * D;
    int64_t head = (int64_t)(((16u - (unsigned)(reinterpret_cast<uintptr_t>(pts) & 15u)) & 15u) / sizeof(T));
    head = head < F ? head : F;
    const int64_t nv = (F - head) / VN;
    const V* __restrict__ vp = reinterpret_cast<const V*>(pts + head);
    T vlo[VN], vhi[VN];
#pragma unroll
    for (int c = 0; c < VN; ++c) { vlo[c] = Lim<T>::inf(); vhi[c] = -Lim<T>::inf(); }
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = gtid; v < nv; v += stride) {
        T e[VN];
        unpack16(vp[v], e);
#pragma unroll
        for (int c = 0; c < VN; ++c) {
            vlo[c] = e[c] < vlo[c] ? e[c] : vlo[c];
            vhi[c] = e[c] > vhi[c] ? e[c] : vhi[c];
        }
    }
    T lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { lo[d] = Lim<T>::inf(); hi[d] = -Lim<T>::inf(); }
#pragma unroll
    for (int c = 0; c < VN; ++c) {
        const int ax = (int)((head + (int64_t)VN * gtid + c) % D);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            lo[d] = (ax == d && vlo[c] < lo[d]) ? vlo[c] : lo[d];
            hi[d] = (ax == d && vhi[c] > hi[d]) ? vhi[c] : hi[d];
        }
    }
    if (gtid == 0) {   // the unaligned head and the tail shorter than one vector
        auto fold = [&](int64_t f) {
            const T v = pts[f];
            const int ax = (int)(f % D);
#pragma unroll
            for (int d = 0; d < D; ++d) {
                lo[d] = (ax == d && v < lo[d]) ? v : lo[d];
                hi[d] = (ax == d && v > hi[d]) ? v : hi[d];
            }
        };
        for (int64_t f = 0; f < head; ++f) fold(f);
        for (int64_t f = head + nv * VN; f < F; ++f) fold(f);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            T a = __shfl_xor_sync(0xffffffffu, lo[d], o), b = __shfl_xor_sync(0xffffffffu, hi[d], o);
            lo[d] = a < lo[d] ? a : lo[d];
            hi[d] = b > hi[d] ? b : hi[d];
        }
    }
    __shared__ T s[BBOX_THREADS / 32][6];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) { s[w][d] = lo[d]; s[w][3 + d] = hi[d]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        T v = s[0][threadIdx.x];
        for (int i = 1; i < BBOX_THREADS / 32; ++i) {
            T u = s[i][threadIdx.x];
            v = threadIdx.x < 3 ? (u < v ? u : v) : (u > v ? u : v);
        }
        partial[blockIdx.x * 6 + threadIdx.x] = v;
    }
}

template <class T>
__global__ void bbox_final_kernel(const T* __restrict__ partial, int nblocks, T* __restrict__ bbox) {
    // 6 warps: warp c reduces component c over all partials
    int c = threadIdx.x >> 5, l = threadIdx.x & 31;
    T v = c < 3 ? Lim<T>::inf() : -Lim<T>::inf();
    for (int i = l; i < nblocks; i += 32) {
        T u = partial[i * 6 + c];
        v = c < 3 ? (u < v ? u : v) : (u > v ? u : v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T u = __shfl_xor_sync(0xffffffffu, v, o);
        v = c < 3 ? (u < v ? u : v) : (u > v ? u : v);
    }
    if (l == 0) bbox[c] = v;
}

template <class T>
void compute_bbox(wtp_ctx* ctx, IndexBuffers& ib, const T* d_pts, int64_t N, int D, double lo[3], double hi[3]) {
    ScopedPhase ph(ctx->timer, PH_BBOX);
    // a multiple of 3 CTAs: the kernel's grid stride must be a multiple of D elements
    int nblocks = (int)std::min<int64_t>((N * D / 4 + BBOX_THREADS - 1) / BBOX_THREADS, (int64_t)kNumSMs * 8);
    nblocks = std::max((nblocks + 2) / 3 * 3, 3);
    T* partial = ib.bbox_partial.as<T>((size_t)nblocks * 6);
    T* bbox = ib.bbox.as<T>(6);
    if (D == 2) bbox_partial_kernel<T, 2><<<nblocks, BBOX_THREADS, 0, ctx->stream>>>(d_pts, N, partial);
    else bbox_partial_kernel<T, 3><<<nblocks, BBOX_THREADS, 0, ctx->stream>>>(d_pts, N, partial);
    LAUNCH_CHECK(ctx);
    bbox_final_kernel<T><<<1, 192, 0, ctx->stream>>>(partial, nblocks, bbox);
    LAUNCH_CHECK(ctx);
    T h[6];
    WTP_CUDA_CHECK(cudaMemcpyAsync(h, bbox, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int d = 0; d < 3; ++d) { lo[d] = d < D ? (double)h[d] : 0.0; hi[d] = d < D ? (double)h[3 + d] : 0.0; }
}
template void compute_bbox<float>(wtp_ctx*, IndexBuffers&, const float*, int64_t, int, double*, double*);
template void compute_bbox<double>(wtp_ctx*, IndexBuffers&, const double*, int64_t, int, double*, double*);

// =========================================================== grid parameters
template <class T>
Grid<T> make_grid(int64_t N, int D, const double lo[3], const double hi[3], double occupancy, double min_cell, int K) {
    Grid<T> g;
    double ext[3] = {0, 0, 0}, vol = 1.0, maxabs = 0.0, maxext = 0.0;
    int deff = 0;
    for (int d = 0; d < D; ++d) {
        // extent in T arithmetic is what cell_coord sees; keep lo exactly as a T
        ext[d] = (double)((T)hi[d] - (T)lo[d]);
        if (!(ext[d] >= 0) || !std::isfinite(ext[d])) throw Error{WTP_ERR_BAD_ARG, "non-finite coordinates in point set"};
        if (ext[d] > 0) { vol *= ext[d]; ++deff; }
        maxabs = std::max(maxabs, std::max(std::fabs(lo[d]), std::fabs(hi[d])));
        maxext = std::max(maxext, ext[d]);
    }
    // Default points per cell: large enough that the K-th neighbour of most queries lies inside the
    // 3^D block (ball of radius ~c holds K + margin points), small enough that the block stays ~10 K.
    if (occupancy <= 0) {
        const double k = K > 0 ? (double)K : 22.0;
        occupancy = std::max(2.0, (deff >= 3 ? 0.36 : deff == 2 ? 0.45 : 0.7) * k);
    }
    double c = 1.0;
    if (deff > 0 && N > 0) c = std::pow(vol * occupancy / (double)N, 1.0 / deff);
    if (!(c > 0) || !std::isfinite(c)) c = maxext > 0 ? maxext : 1.0;
    // keep the cell size representable relative to the extent (at most 2^20 cells per dimension)
    c = std::max(c, maxext / 1048576.0);
    if (min_cell > 0) c = std::max(c, min_cell);
    const double cap = std::min(1073741824.0, std::max(4096.0, 4.0 * (double)N));
    for (;;) {
        double prod = 1.0;
        for (int d = 0; d < 3; ++d) {
            double nd = d < D ? std::floor(ext[d] / c) + 1.0 : 1.0;
            g.n[d] = (int)std::min(nd, 1048576.0);
            prod *= g.n[d];
        }
        if (prod <= cap) { g.ncells = (uint32_t)prod; break; }
        c *= 1.1;
    }
    for (int d = 0; d < 3; ++d) g.lo[d] = d < D ? (T)lo[d] : (T)0;
    g.c = (T)c;
    g.inv_c = (T)1 / g.c;
    const double eps = sizeof(T) == 4 ? (double)FLT_EPSILON : DBL_EPSILON;
    // cell boundaries lo + i*c and the (v-lo)*inv_c rounding are each off by a few ulps of
    // the coordinate magnitude; 16 ulps of slack keeps every pruning bound conservative.
    g.slack = (T)(16.0 * eps * std::max(maxabs, maxext) + 4.0 * eps * c);
    g.w_lo = 0;
    g.w_hi = g.n[D - 1] - 1;
    return g;
}
template Grid<float> make_grid<float>(int64_t, int, const double*, const double*, double, double, int);
template Grid<double> make_grid<double>(int64_t, int, const double*, const double*, double, double, int);

// ================================================================ cell keys
template <class T, int D>
__global__ void __launch_bounds__(256) cellkey_kernel(const T* __restrict__ pts, int64_t N, Grid<T> g,
                                                      uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int cx = cell_coord(g, pts[i * D + 0], 0);
    int cy = cell_coord(g, pts[i * D + 1], 1);
    int cz = D == 3 ? cell_coord(g, pts[i * D + (D - 1)], 2) : 0;
    keys[i] = ((uint32_t)cz * (uint32_t)g.n[1] + (uint32_t)cy) * (uint32_t)g.n[0] + (uint32_t)cx;
    vals[i] = (uint32_t)i;
}

// =============================================================== radix sort
// Stable LSD radix sort of (cell key, point id) pairs, one byte per pass. Per pass:
//   rs_hist_kernel    per-tile digit histogram            (reads 4 B/item)
//   exclusive scan    over the digit-major histogram
//   rs_scatter_kernel stable in-tile ranking in shared memory, coalesced run writes
//                                                          (reads 8 B, writes 8 B/item)
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 4096 keys per CTA
constexpr int RS_WARP_ITEMS = RS_TILE / RS_WARPS;  // 512 consecutive keys per warp

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n, int shift,
                                                             uint32_t* __restrict__ block_hist, int num_blocks) {
    // one private histogram per warp (shared-memory atomics, a few-way conflicts at most), 16-byte key loads
    __shared__ uint32_t h[RS_WARPS][256];
    const int w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
    if (base + RS_TILE <= n) {
        const uint4* k4 = reinterpret_cast<const uint4*>(keys + base);     // RS_TILE is a multiple of 4: 16-byte aligned
#pragma unroll
        for (int r = 0; r < RS_ITEMS / 4; ++r) {
            const uint4 v = k4[r * RS_THREADS + threadIdx.x];
            atomicAdd(&h[w][(v.x >> shift) & 255u], 1u);
            atomicAdd(&h[w][(v.y >> shift) & 255u], 1u);
            atomicAdd(&h[w][(v.z >> shift) & 255u], 1u);
            atomicAdd(&h[w][(v.w >> shift) & 255u], 1u);
        }
    } else {
        for (uint32_t i = base + threadIdx.x; i < n; i += RS_THREADS) atomicAdd(&h[w][(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    uint32_t tot = 0;
#pragma unroll
    for (int i = 0; i < RS_WARPS; ++i) tot += h[i][threadIdx.x];
    block_hist[threadIdx.x * num_blocks + blockIdx.x] = tot;
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                uint32_t n, int shift, const uint32_t* __restrict__ block_offsets,
                                                                int num_blocks) {
    __shared__ uint32_t warp_cnt[RS_WARPS][256];
    __shared__ uint32_t digit_base[256];
    __shared__ uint32_t global_base[256];
    __shared__ uint32_t s_warp_tot[RS_WARPS];
    __shared__ uint32_t s_keys[RS_TILE];
    __shared__ uint32_t s_vals[RS_TILE];

    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const uint32_t tile_base = blockIdx.x * RS_TILE;
    const uint32_t n_valid = min((uint32_t)RS_TILE, n - tile_base);
    for (int i = tid; i < RS_WARPS * 256; i += RS_THREADS) (&warp_cnt[0][0])[i] = 0;
    __syncthreads();

    uint32_t key[RS_ITEMS];
    uint16_t rank[RS_ITEMS];
    // Warp w owns the 512 consecutive keys [w*512, (w+1)*512) of the tile, walked in 16
    // rounds of 32: tile order = (warp, round, lane) = input order, which keeps the sort stable.
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        uint32_t local = w * RS_WARP_ITEMS + r * 32 + lane;
        bool valid = local < n_valid;
        key[r] = valid ? keys_in[tile_base + local] : 0xffffffffu;
        // invalid tail entries sort as digit 255 behind every valid key of the tile
        uint32_t digit = valid ? ((key[r] >> shift) & 255u) : 255u;
        unsigned peers = __match_any_sync(0xffffffffu, digit);
        uint32_t before = warp_cnt[w][digit];
        rank[r] = (uint16_t)(before + __popc(peers & ((1u << lane) - 1)));
        __syncwarp();
        if ((peers & ((1u << lane) - 1)) == 0) warp_cnt[w][digit] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // thread d: exclusive scan over the warps of digit d, then exclusive scan over digits
    uint32_t total = 0;
    {
        const int d = tid;
#pragma unroll
        for (int i = 0; i < RS_WARPS; ++i) {
            uint32_t c = warp_cnt[i][d];
            warp_cnt[i][d] = total;
            total += c;
        }
        uint32_t incl = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp_tot[w] = incl;
        __syncthreads();
        uint32_t wbase = 0;
        for (int i = 0; i < w; ++i) wbase += s_warp_tot[i];
        uint32_t excl = wbase + incl - total;
        digit_base[d] = excl;
        global_base[d] = block_offsets[d * num_blocks + blockIdx.x] - excl;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        uint32_t local = w * RS_WARP_ITEMS + r * 32 + lane;
        bool valid = local < n_valid;
        uint32_t digit = valid ? ((key[r] >> shift) & 255u) : 255u;
        uint32_t lp = digit_base[digit] + warp_cnt[w][digit] + rank[r];
        s_keys[lp] = key[r];
        s_vals[lp] = valid ? vals_in[tile_base + local] : 0u;
    }
    __syncthreads();
    for (uint32_t i = tid; i < n_valid; i += RS_THREADS) {
        uint32_t k = s_keys[i];
        uint32_t out = global_base[(k >> shift) & 255u] + i;
        keys_out[out] = k;
        vals_out[out] = s_vals[i];
    }
}

// ==================================================================== scans
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

template <class A>
__device__ inline A block_exclusive_scan(A v, A* s_warp /*[SC_THREADS/32]*/, A* block_total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    A incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        A t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    A base = 0, tot = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
        A c = s_warp[i];
        if (i < w) base += c;
        tot += c;
    }
    if (block_total) *block_total = tot;
    return base + incl - v;
}

template <class A>
__global__ void __launch_bounds__(SC_THREADS) scan_reduce_kernel(const uint32_t* __restrict__ in, int64_t n, A* __restrict__ block_sums) {
    __shared__ A s_warp[SC_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    A v = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) if (base + i < n) v += (A)in[base + i];
    A tot;
    (void)block_exclusive_scan<A>(v, s_warp, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

template <class A>
__global__ void __launch_bounds__(1024) scan_sums_kernel(A* __restrict__ block_sums, int64_t nb) {
    // single CTA: in-place exclusive scan of nb block sums, chunk by chunk with a carry;
    // block_sums[nb] receives the grand total.
    __shared__ A s_warp[32];
    __shared__ A s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < nb; c0 += blockDim.x) {
        int64_t i = c0 + threadIdx.x;
        A v = i < nb ? block_sums[i] : (A)0;
        A tot;
        A ex = block_exclusive_scan<A>(v, s_warp, &tot);
        A carry = s_carry;
        if (i < nb) block_sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[nb] = s_carry;
}

template <class A>
__global__ void __launch_bounds__(SC_THREADS) scan_apply_kernel(const uint32_t* in, A* out, int64_t n,
                                                                const A* __restrict__ block_sums, int64_t nb) {
    // in and out may alias (in-place scan of the radix histogram): every thread reads all
    // of its own items before it writes them, and touches nobody else's.
    __shared__ A s_warp[SC_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    A item[SC_ITEMS];
    A v = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) { item[i] = base + i < n ? (A)in[base + i] : (A)0; v += item[i]; }
    A ex = block_exclusive_scan<A>(v, s_warp, nullptr) + block_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += item[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = block_sums[nb];
}

template <class A>
static void exclusive_scan_impl(wtp_ctx* ctx, DevBuf& tmp, const uint32_t* d_in, A* d_out, int64_t n) {
    if (n <= 0) { WTP_CUDA_CHECK(cudaMemsetAsync(d_out, 0, sizeof(A), ctx->stream)); return; }
    int64_t nb = (n + SC_TILE - 1) / SC_TILE;
    A* sums = tmp.as<A>((size_t)nb + 1);
    scan_reduce_kernel<A><<<(unsigned)nb, SC_THREADS, 0, ctx->stream>>>(d_in, n, sums);
    LAUNCH_CHECK(ctx);
    scan_sums_kernel<A><<<1, 1024, 0, ctx->stream>>>(sums, nb);
    LAUNCH_CHECK(ctx);
    scan_apply_kernel<A><<<(unsigned)nb, SC_THREADS, 0, ctx->stream>>>(d_in, d_out, n, sums, nb);
    LAUNCH_CHECK(ctx);
}
void exclusive_scan_u32(wtp_ctx* ctx, DevBuf& tmp, const uint32_t* d_in, uint32_t* d_out, int64_t n) {
    exclusive_scan_impl<uint32_t>(ctx, tmp, d_in, d_out, n);
}
void exclusive_scan_u32_to_i64(wtp_ctx* ctx, DevBuf& tmp, const uint32_t* d_in, int64_t* d_out, int64_t n) {
    exclusive_scan_impl<unsigned long long>(ctx, tmp, d_in, reinterpret_cast<unsigned long long*>(d_out), n);
}

// Sorts the (key, value) pairs held in ib.keys_a / ib.vals_a by the low `bits` bits of the
// key; on return ib.keys_a / ib.vals_a hold the sorted pairs. Returns the passes run.
int radix_sort_pairs(wtp_ctx* ctx, IndexBuffers& ib, int64_t N, int bits) {
    ScopedPhase ph(ctx->timer, PH_SORT);
    cudaStream_t st = ctx->stream;
    uint32_t* keys_a = ib.keys_a.get<uint32_t>();
    uint32_t* vals_a = ib.vals_a.get<uint32_t>();
    uint32_t* keys_b = ib.keys_b.as<uint32_t>((size_t)N);
    uint32_t* vals_b = ib.vals_b.as<uint32_t>((size_t)N);
    const int passes = (bits + 7) / 8;
    const int num_blocks = (int)((N + RS_TILE - 1) / RS_TILE);
    uint32_t* hist = ib.block_hist.as<uint32_t>((size_t)256 * num_blocks + 1);
    for (int p = 0; p < passes; ++p) {
        rs_hist_kernel<<<num_blocks, RS_THREADS, 0, st>>>(keys_a, (uint32_t)N, 8 * p, hist, num_blocks);
        LAUNCH_CHECK(ctx);
        exclusive_scan_u32(ctx, ib.scan_tmp, hist, hist, (int64_t)256 * num_blocks);
        rs_scatter_kernel<<<num_blocks, RS_THREADS, 0, st>>>(keys_a, vals_a, keys_b, vals_b, (uint32_t)N, 8 * p, hist, num_blocks);
        LAUNCH_CHECK(ctx);
        std::swap(keys_a, keys_b);
        std::swap(vals_a, vals_b);
    }
    if (passes & 1) {
        std::swap(ib.keys_a.p, ib.keys_b.p); std::swap(ib.keys_a.cap, ib.keys_b.cap);
        std::swap(ib.vals_a.p, ib.vals_b.p); std::swap(ib.vals_a.cap, ib.vals_b.cap);
    }
    return passes;
}

// ====================================================== gather + cell starts
template <class T, int D>
__global__ void __launch_bounds__(256) reorder_kernel(const T* __restrict__ pts, const uint32_t* __restrict__ keys,
                                                      const uint32_t* __restrict__ vals, uint32_t N, uint32_t ncells,
                                                      P4<T>* __restrict__ sorted, uint32_t* __restrict__ cell_start) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    uint32_t i = vals[j];
    P4<T> p;
    p.x = pts[(size_t)i * D + 0];
    p.y = pts[(size_t)i * D + 1];
    p.z = D == 3 ? pts[(size_t)i * D + (D - 1)] : (T)0;
    p.w = idx_bits((T)0, i);
    sorted[j] = p;
    uint32_t key = keys[j];
    if (j == 0) {
        for (uint32_t c = 0; c <= key; ++c) cell_start[c] = 0;
    } else {
        uint32_t prev = keys[j - 1];
        for (uint32_t c = prev + 1; c <= key; ++c) cell_start[c] = j;  // empty cells in the gap start here too
    }
    if (j == N - 1)
        for (uint32_t c = key + 1; c <= ncells; ++c) cell_start[c] = N;
}


// ====================================================== counting-sort build
// The cell keys are small integers and the index needs them grouped, not globally merged: count the points of every
// cell while computing the keys (the atomic's return value is the point's arrival number in its cell), scan the counts
// into cell_start, store every point's caller index at cell_start[key] + arrival number (a 4 N-byte array that stays in
// L2), then put the (at most a few dozen) entries of every cell in ascending caller index — the order the stable radix
// sort gives, so both builds produce the same records — while gathering the records. No digit passes.
//   cell_count_kernel        keys, arrival numbers, per-cell counts (fused with the key computation)
//   count_keys_kernel        the same from existing keys (windowed build)
//   cell_scatter_kernel      caller index -> order[cell_start[key] + arrival number]
//   cell_place_kernel        one thread per entry: rank among the entries of its cell by caller index, gather, store
//   cell_place_big_kernel    cells above CS_RANK_LIMIT points: one CTA per cell, bitonic sort in shared memory (in global
//                            memory above CS_BIG_MAX points: degenerate clouds only)
constexpr uint32_t CS_RANK_LIMIT = 64;      // cells up to this many points are ranked by counting (O(n) per point)
constexpr uint32_t CS_BIG_MAX = 8192;       // heaviest cell sorted in shared memory

template <class T, int D>
__device__ inline uint32_t cell_key_of(const Grid<T>& g, T x, T y, T z) {
    const int cx = cell_coord(g, x, 0);
    const int cy = cell_coord(g, y, 1);
    const int cz = D == 3 ? cell_coord(g, z, 2) : 0;
    return ((uint32_t)cz * (uint32_t)g.n[1] + (uint32_t)cy) * (uint32_t)g.n[0] + (uint32_t)cx;
}

template <class T, int D>
__global__ void __launch_bounds__(256) cell_count_kernel(const T* __restrict__ pts, int64_t N, Grid<T> g, uint32_t* __restrict__ keys,
                                                         uint32_t* __restrict__ arrival, uint32_t* __restrict__ cell_count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint32_t key = cell_key_of<T, D>(g, pts[i * D + 0], pts[i * D + 1], pts[i * D + (D - 1)]);
    keys[i] = key;
    arrival[i] = atomicAdd(cell_count + key, 1u);
}

__global__ void __launch_bounds__(256) count_keys_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ arrival,
                                                         uint32_t* __restrict__ cell_count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) arrival[i] = atomicAdd(cell_count + keys[i], 1u);
}

template <class T, int D>
__device__ inline P4<T> point_record(const T* __restrict__ pts, uint32_t i) {
    P4<T> p;
    p.x = pts[(size_t)i * D + 0];
    p.y = pts[(size_t)i * D + 1];
    p.z = D == 3 ? pts[(size_t)i * D + (D - 1)] : (T)0;
    p.w = idx_bits((T)0, i);
    return p;
}

__global__ void __launch_bounds__(256) cell_scatter_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ arrival,
                                                           const uint32_t* __restrict__ ids, uint32_t n, const uint32_t* __restrict__ cell_start,
                                                           uint32_t* __restrict__ order) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) order[cell_start[keys[i]] + arrival[i]] = ids ? ids[i] : i;
}

template <class T, int D>
__global__ void __launch_bounds__(256) cell_place_kernel(const T* __restrict__ pts, const uint32_t* __restrict__ order, uint32_t n, Grid<T> g,
                                                         uint32_t key_lo, const uint32_t* __restrict__ cell_start, P4<T>* __restrict__ sorted,
                                                         uint32_t* __restrict__ big_list) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t mine = order[j];
    const P4<T> me = point_record<T, D>(pts, mine);
    const uint32_t key = cell_key_of<T, D>(g, me.x, me.y, me.z) - key_lo;
    const uint32_t s = cell_start[key], e = cell_start[key + 1];
    if (e - s > CS_RANK_LIMIT) {
        if (j == s) big_list[1 + atomicAdd(big_list, 1u)] = key;
        return;
    }
    uint32_t rank = 0;
    for (uint32_t m = s; m < e; ++m) rank += order[m] < mine ? 1u : 0u;
    sorted[s + rank] = me;
}

// Ascending sort of a[0 .. n) by the threads of one CTA: the bitonic network in its all-ascending form (the first stage
// of every merge compares i with its mirror image in the run, the later ones i with i + j), so that the positions
// n .. 2^ceil(log2 n) behave as +inf without existing: a comparison that would touch them changes nothing.
__device__ inline void cta_sort_u32(uint32_t* a, uint32_t n) {
    uint32_t np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (uint32_t k = 2; k <= np2; k <<= 1) {
        const uint32_t hk = k >> 1;
        for (uint32_t t = threadIdx.x; t < (np2 >> 1); t += blockDim.x) {
            const uint32_t r = t & (hk - 1), i = ((t - r) << 1) + r, l = i + k - 1 - 2 * r;      // mirror image of i in its run of k
            if (l < n) { const uint32_t x = a[i], y = a[l]; if (x > y) { a[i] = y; a[l] = x; } }
        }
        __syncthreads();
        for (uint32_t j = hk >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < (np2 >> 1); t += blockDim.x) {
                const uint32_t i = 2 * t - (t & (j - 1)), l = i + j;
                if (l < n) { const uint32_t x = a[i], y = a[l]; if (x > y) { a[i] = y; a[l] = x; } }
            }
            __syncthreads();
        }
    }
}

template <class T, int D>
__global__ void __launch_bounds__(256) cell_place_big_kernel(const T* __restrict__ pts, uint32_t* __restrict__ order,
                                                             const uint32_t* __restrict__ cell_start, P4<T>* __restrict__ sorted,
                                                             const uint32_t* __restrict__ big_list) {
    __shared__ uint32_t s_idx[CS_BIG_MAX];
    const uint32_t n_big = big_list[0];
    for (uint32_t b = blockIdx.x; b < n_big; b += gridDim.x) {
        const uint32_t cell = big_list[1 + b];
        const uint32_t s = cell_start[cell], n = cell_start[cell + 1] - s;
        uint32_t* a = order + s;                              // cells above the shared-memory size are sorted where they are
        if (n <= CS_BIG_MAX) {
            for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) s_idx[t] = order[s + t];
            a = s_idx;
        }
        __syncthreads();
        cta_sort_u32(a, n);
        for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) sorted[s + t] = point_record<T, D>(pts, a[t]);
        __syncthreads();
    }
}

// Groups the `n` points (keys[i], arrival[i], ids ? ids[i] : i) by key into `sorted`, ascending caller index within
// a cell, and fills cell_start[0 .. ncells]. cell_count[0 .. ncells) holds the points per cell (keys are cell ids
// minus key_lo). No host round trip: the placement kernel lists the cells too heavy for it and the kernel behind it
// takes whatever is on the list.
template <class T>
static void place_counted(wtp_ctx* ctx, IndexBuffers& ib, const T* d_pts, int D, const Grid<T>& g, uint32_t key_lo, const uint32_t* keys,
                          const uint32_t* arrival, const uint32_t* ids, uint32_t n, uint32_t ncells, uint32_t* cell_count, P4<T>* sorted,
                          uint32_t* cell_start) {
    cudaStream_t st = ctx->stream;
    const unsigned nb = (n + 255u) / 256u;
    uint32_t* order = ib.staging.as<uint32_t>((size_t)n);
    uint32_t* big_list = ib.keys_b.as<uint32_t>((size_t)n / CS_RANK_LIMIT + 2);
    {
        ScopedPhase ph(ctx->timer, PH_SORT);
        exclusive_scan_u32(ctx, ib.scan_tmp, cell_count, cell_start, (int64_t)ncells);
        WTP_CUDA_CHECK(cudaMemsetAsync(big_list, 0, sizeof(uint32_t), st));
        cell_scatter_kernel<<<nb, 256, 0, st>>>(keys, arrival, ids, n, cell_start, order);
        LAUNCH_CHECK(ctx);
    }
    {
        ScopedPhase ph(ctx->timer, PH_REORDER);
        if (D == 2) cell_place_kernel<T, 2><<<nb, 256, 0, st>>>(d_pts, order, n, g, key_lo, cell_start, sorted, big_list);
        else cell_place_kernel<T, 3><<<nb, 256, 0, st>>>(d_pts, order, n, g, key_lo, cell_start, sorted, big_list);
        LAUNCH_CHECK(ctx);
        if (D == 2) cell_place_big_kernel<T, 2><<<kNumSMs * 4, 256, 0, st>>>(d_pts, order, cell_start, sorted, big_list);
        else cell_place_big_kernel<T, 3><<<kNumSMs * 4, 256, 0, st>>>(d_pts, order, cell_start, sorted, big_list);
        LAUNCH_CHECK(ctx);
    }
}

// The counting build pays off while the 4 n-byte `order` array (and the cell starts) stay in the 126 MB L2: its
// scatter is n random 4-byte writes. Measured: 10 M points 0.40 ms against 0.55 ms for the radix build; 100 M points
// 8.1 ms against ~5.5 ms (the scatter alone 4.5 ms once every write is a DRAM read-modify-write). Above 64 MB of
// `order` the radix build runs. WTP_RADIX_BUILD=1 forces it (read per build: the tests switch it inside one process).
static bool counting_build_enabled(size_t n) {
    const char* e = std::getenv("WTP_RADIX_BUILD");
    return !(e && e[0] == '1') && n * sizeof(uint32_t) <= ((size_t)64 << 20);
}

template <class T>
int build_index(wtp_ctx* ctx, IndexBuffers& ib, const T* d_pts, int64_t N, int D, const Grid<T>& g) {
    WTP_REQUIRE(N > 0 && N < (int64_t)0xfffffff0u, WTP_ERR_BAD_ARG, "point count must be in [1, 2^32)");
    cudaStream_t st = ctx->stream;
    uint32_t* keys_a = ib.keys_a.as<uint32_t>((size_t)N);
    uint32_t* vals_a = ib.vals_a.as<uint32_t>((size_t)N);
    P4<T>* sorted = ib.sorted.as<P4<T>>((size_t)N);
    uint32_t* cell_start = ib.cell_start.as<uint32_t>((size_t)g.ncells + 1);
    ib.cs_rebase = 0;
    const unsigned nb256 = (unsigned)((N + 255) / 256);
    if (counting_build_enabled((size_t)N)) {
        uint32_t* cell_count = ib.block_hist.as<uint32_t>((size_t)g.ncells + 2);
        {
            ScopedPhase ph(ctx->timer, PH_CELLKEY);
            WTP_CUDA_CHECK(cudaMemsetAsync(cell_count, 0, ((size_t)g.ncells + 2) * sizeof(uint32_t), st));
            uint32_t* arrival = ib.vals_b.as<uint32_t>((size_t)N);
            if (D == 2) cell_count_kernel<T, 2><<<nb256, 256, 0, st>>>(d_pts, N, g, keys_a, arrival, cell_count);
            else cell_count_kernel<T, 3><<<nb256, 256, 0, st>>>(d_pts, N, g, keys_a, arrival, cell_count);
            LAUNCH_CHECK(ctx);
        }
        place_counted<T>(ctx, ib, d_pts, D, g, 0u, keys_a, ib.vals_b.get<uint32_t>(), nullptr, (uint32_t)N, g.ncells, cell_count, sorted, cell_start);
        return 0;
    }
    {
        ScopedPhase ph(ctx->timer, PH_CELLKEY);
        if (D == 2) cellkey_kernel<T, 2><<<nb256, 256, 0, st>>>(d_pts, N, g, keys_a, vals_a);
        else cellkey_kernel<T, 3><<<nb256, 256, 0, st>>>(d_pts, N, g, keys_a, vals_a);
        LAUNCH_CHECK(ctx);
    }
    int bits = 0;
    while (bits < 32 && ((uint64_t)1 << bits) < (uint64_t)g.ncells) ++bits;
    const int passes = radix_sort_pairs(ctx, ib, N, bits);
    keys_a = ib.keys_a.get<uint32_t>();
    vals_a = ib.vals_a.get<uint32_t>();
    {
        ScopedPhase ph(ctx->timer, PH_REORDER);
        if (D == 2) reorder_kernel<T, 2><<<nb256, 256, 0, st>>>(d_pts, keys_a, vals_a, (uint32_t)N, g.ncells, sorted, cell_start);
        else reorder_kernel<T, 3><<<nb256, 256, 0, st>>>(d_pts, keys_a, vals_a, (uint32_t)N, g.ncells, sorted, cell_start);
        LAUNCH_CHECK(ctx);
    }
    return passes;
}
template int build_index<float>(wtp_ctx*, IndexBuffers&, const float*, int64_t, int, const Grid<float>&);
template int build_index<double>(wtp_ctx*, IndexBuffers&, const double*, int64_t, int, const Grid<double>&);

// =========================================================== windowed index
// A sharded k-NN rank answers the sorted positions [sb, se) only. Its queries live in a run of x-rows of cells, so
// their neighbourhoods live in a few layers of the slowest axis: instead of sorting all N points, the rank
//   1. computes every cell key and counts the points of every layer of the slowest axis;
//   2. scans the counts, finds the layers of positions sb and se-1, adds `halo` layers on either side (the
//      window), and learns how many points lie below and inside it (one small D2H);
//   3. compacts the (key, id) pairs of the window in input order (stable), sorts them, and gathers the records.
// The sorted records of the window are exactly the run [P0, P0 + M) of the whole sorted order (same stable sort on
// the same keys), so rows, their order and wtp_shard_owned are identical to the replicated build.
constexpr int WC_THREADS = 256;
constexpr int WC_ITEMS = 16;
constexpr int WC_TILE = WC_THREADS * WC_ITEMS;

constexpr int WH_MAX_LAYERS = 8192;     // layer histogram in shared memory (32 KB)
constexpr int WH_THREADS = 256;

// every cell key + the number of points per layer of the slowest axis: per-CTA histogram in shared memory,
// flushed with one global atomic per non-empty bin
template <class T, int D>
__global__ void __launch_bounds__(WH_THREADS) cellkey_hist_kernel(const T* __restrict__ pts, int64_t N, Grid<T> g, uint32_t* __restrict__ keys,
                                                                  uint32_t* __restrict__ layer_hist) {
    extern __shared__ uint32_t s_hist[];
    const int n_layers = g.n[D - 1];
    for (int i = threadIdx.x; i < n_layers; i += WH_THREADS) s_hist[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * WH_THREADS + threadIdx.x; i < N; i += (int64_t)gridDim.x * WH_THREADS) {
        const int cx = cell_coord(g, pts[i * D + 0], 0);
        const int cy = cell_coord(g, pts[i * D + 1], 1);
        const int cz = D == 3 ? cell_coord(g, pts[i * D + (D - 1)], 2) : 0;
        keys[i] = ((uint32_t)cz * (uint32_t)g.n[1] + (uint32_t)cy) * (uint32_t)g.n[0] + (uint32_t)cx;
        atomicAdd(&s_hist[D == 3 ? cz : cy], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_layers; i += WH_THREADS) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(layer_hist + i, c);
    }
}

// One CTA: exclusive scan of the layer counts in shared memory, the layers of the sorted positions sb and se-1,
// the window. out: {P0, M, w_lo, w_hi}
__global__ void __launch_bounds__(1024) window_pick_kernel(const uint32_t* __restrict__ layer_hist, uint32_t n_layers, uint32_t sb, uint32_t se,
                                                           uint32_t halo, uint32_t* __restrict__ out) {
    __shared__ uint32_t s_pre[WH_MAX_LAYERS + 1];
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry, s_first, s_last;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t c0 = 0; c0 < n_layers; c0 += 1024) {
        const uint32_t i = c0 + threadIdx.x;
        const uint32_t v = i < n_layers ? layer_hist[i] : 0u;
        uint32_t incl = v;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        uint32_t base = s_carry;
        for (int k = 0; k < w; ++k) base += s_warp[k];
        if (i < n_layers) s_pre[i] = base + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = base + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) s_pre[n_layers] = s_carry;
    __syncthreads();
    // the layer of position p: the one with pre[l] <= p < pre[l + 1] (never an empty layer)
    for (uint32_t l = threadIdx.x; l < n_layers; l += 1024) {
        if (s_pre[l] <= sb && sb < s_pre[l + 1]) s_first = l;
        if (s_pre[l] <= se - 1 && se - 1 < s_pre[l + 1]) s_last = l;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t w_lo = s_first > halo ? s_first - halo : 0u;
        const uint32_t w_hi = s_last + halo < n_layers - 1 ? s_last + halo : n_layers - 1;
        out[0] = s_pre[w_lo];
        out[1] = s_pre[w_hi + 1] - s_pre[w_lo];
        out[2] = w_lo;
        out[3] = w_hi;
    }
}

__global__ void __launch_bounds__(WC_THREADS) window_count_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t key_lo, uint32_t key_hi,
                                                                  uint32_t* __restrict__ block_counts) {
    const uint32_t base = blockIdx.x * WC_TILE;
    uint32_t c = 0;
    for (uint32_t i = base + threadIdx.x; i < min(n, base + WC_TILE); i += WC_THREADS) {
        const uint32_t k = keys[i];
        c += (k >= key_lo && k < key_hi) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __shared__ uint32_t s[WC_THREADS / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < WC_THREADS / 32; ++i) t += s[i];
        block_counts[blockIdx.x] = t;
    }
}

// stable: thread t of a CTA takes the WC_ITEMS consecutive keys [t * WC_ITEMS, ...) of the CTA's tile
__global__ void __launch_bounds__(WC_THREADS) window_compact_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t key_lo, uint32_t key_hi,
                                                                    const uint32_t* __restrict__ block_offsets,
                                                                    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t s_warp[WC_THREADS / 32];
    const uint32_t first = blockIdx.x * WC_TILE + threadIdx.x * WC_ITEMS;
    uint32_t k[WC_ITEMS];
    uint32_t c = 0;
    if (first + WC_ITEMS <= n) {
        const uint4* k4 = reinterpret_cast<const uint4*>(keys + first);
#pragma unroll
        for (int r = 0; r < WC_ITEMS / 4; ++r) { const uint4 v = k4[r]; k[4 * r] = v.x; k[4 * r + 1] = v.y; k[4 * r + 2] = v.z; k[4 * r + 3] = v.w; }
    } else {
#pragma unroll
        for (int r = 0; r < WC_ITEMS; ++r) k[r] = first + r < n ? keys[first + r] : 0xffffffffu;   // key_hi <= ncells < 2^32 - 1
    }
#pragma unroll
    for (int r = 0; r < WC_ITEMS; ++r) c += (k[r] >= key_lo && k[r] < key_hi) ? 1u : 0u;
    uint32_t out = block_exclusive_scan<uint32_t>(c, s_warp, nullptr) + block_offsets[blockIdx.x];
#pragma unroll
    for (int r = 0; r < WC_ITEMS; ++r) {
        if (k[r] >= key_lo && k[r] < key_hi) {
            keys_out[out] = k[r] - key_lo;
            vals_out[out] = first + r;
            ++out;
        }
    }
}

template <class T>
bool build_index_window(wtp_ctx* ctx, IndexBuffers& ib, const T* d_pts, int64_t N, int D, Grid<T>& g, int64_t sb, int64_t se, int halo,
                        IndexWindow* win, int* passes_out) {
    WTP_REQUIRE(N > 0 && N < (int64_t)0xfffffff0u, WTP_ERR_BAD_ARG, "point count must be in [1, 2^32)");
    const uint32_t rows_per_layer = D == 3 ? (uint32_t)g.n[1] : 1u;
    const uint32_t n_layers = (uint32_t)g.n[D - 1];
    if (se <= sb || n_layers > (uint32_t)WH_MAX_LAYERS || (int64_t)n_layers <= 2 * halo + 1) return false;   // nothing to gain: build the whole index
    cudaStream_t st = ctx->stream;
    uint32_t* keys_all = ib.keys_b.as<uint32_t>((size_t)N);
    uint32_t* hist = ib.unit_hist.as<uint32_t>((size_t)n_layers + 8);
    uint32_t* pick = hist + n_layers;
    {
        ScopedPhase ph(ctx->timer, PH_CELLKEY);
        WTP_CUDA_CHECK(cudaMemsetAsync(hist, 0, ((size_t)n_layers + 8) * sizeof(uint32_t), st));
        const unsigned nbh = (unsigned)std::min<int64_t>((N + WH_THREADS - 1) / WH_THREADS, (int64_t)kNumSMs * 8);
        const size_t smem = (size_t)n_layers * sizeof(uint32_t);
        if (D == 2) cellkey_hist_kernel<T, 2><<<nbh, WH_THREADS, smem, st>>>(d_pts, N, g, keys_all, hist);
        else cellkey_hist_kernel<T, 3><<<nbh, WH_THREADS, smem, st>>>(d_pts, N, g, keys_all, hist);
        LAUNCH_CHECK(ctx);
        window_pick_kernel<<<1, 1024, 0, st>>>(hist, n_layers, (uint32_t)sb, (uint32_t)se, (uint32_t)halo, pick);
        LAUNCH_CHECK(ctx);
        uint32_t* h = reinterpret_cast<uint32_t*>(static_cast<char*>(ctx->h_pinned) + 2048);
        WTP_CUDA_CHECK(cudaMemcpyAsync(h, pick, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        WTP_CUDA_CHECK(cudaStreamSynchronize(st));
        win->P0 = h[0]; win->M = h[1]; win->w_lo = (int)h[2]; win->w_hi = (int)h[3];
    }
    win->key_lo = (uint32_t)win->w_lo * rows_per_layer * (uint32_t)g.n[0];
    win->key_hi = ((uint32_t)win->w_hi + 1u) * rows_per_layer * (uint32_t)g.n[0];
    const int64_t M = win->M;
    WTP_REQUIRE(M > 0 && win->P0 <= sb && se <= win->P0 + M, WTP_ERR_CUDA, "windowed index: the window does not hold the owned run");
    uint32_t* keys_a = ib.keys_a.as<uint32_t>((size_t)N);
    uint32_t* vals_a = ib.vals_a.as<uint32_t>((size_t)N);
    {
        ScopedPhase ph(ctx->timer, PH_CELLKEY);
        const int64_t nb = (N + WC_TILE - 1) / WC_TILE;
        uint32_t* counts = ib.block_hist.as<uint32_t>((size_t)nb + 1);
        window_count_kernel<<<(unsigned)nb, WC_THREADS, 0, st>>>(keys_all, (uint32_t)N, win->key_lo, win->key_hi, counts);
        LAUNCH_CHECK(ctx);
        exclusive_scan_u32(ctx, ib.scan_tmp, counts, counts, nb);
        window_compact_kernel<<<(unsigned)nb, WC_THREADS, 0, st>>>(keys_all, (uint32_t)N, win->key_lo, win->key_hi, counts, keys_a, vals_a);
        LAUNCH_CHECK(ctx);
    }
    const uint32_t ncells_w = win->key_hi - win->key_lo;
    P4<T>* sorted = ib.sorted.as<P4<T>>((size_t)M);
    uint32_t* cell_start = ib.cell_start.as<uint32_t>((size_t)ncells_w + 1);
    ib.cs_rebase = (int64_t)win->key_lo;
    g.w_lo = win->w_lo;
    g.w_hi = win->w_hi;
    if (counting_build_enabled((size_t)M)) {
        uint32_t* cell_count = ib.block_hist.as<uint32_t>((size_t)ncells_w + 2);     // the compaction's block counts are consumed
        {
            ScopedPhase ph(ctx->timer, PH_SORT);
            WTP_CUDA_CHECK(cudaMemsetAsync(cell_count, 0, ((size_t)ncells_w + 2) * sizeof(uint32_t), st));
            count_keys_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(keys_a, (uint32_t)M, ib.vals_b.as<uint32_t>((size_t)M), cell_count);
            LAUNCH_CHECK(ctx);
        }
        place_counted<T>(ctx, ib, d_pts, D, g, win->key_lo, keys_a, ib.vals_b.get<uint32_t>(), vals_a, (uint32_t)M, ncells_w, cell_count, sorted,
                         cell_start);
        if (passes_out) *passes_out = 0;
        return true;
    }
    int bits = 0;
    while (bits < 32 && ((uint64_t)1 << bits) < (uint64_t)ncells_w) ++bits;
    const int passes = radix_sort_pairs(ctx, ib, M, bits);
    keys_a = ib.keys_a.get<uint32_t>();
    vals_a = ib.vals_a.get<uint32_t>();
    {
        ScopedPhase ph(ctx->timer, PH_REORDER);
        const unsigned nbm = (unsigned)((M + 255) / 256);
        if (D == 2) reorder_kernel<T, 2><<<nbm, 256, 0, st>>>(d_pts, keys_a, vals_a, (uint32_t)M, ncells_w, sorted, cell_start);
        else reorder_kernel<T, 3><<<nbm, 256, 0, st>>>(d_pts, keys_a, vals_a, (uint32_t)M, ncells_w, sorted, cell_start);
        LAUNCH_CHECK(ctx);
    }
    if (passes_out) *passes_out = passes;
    return true;
}
template bool build_index_window<float>(wtp_ctx*, IndexBuffers&, const float*, int64_t, int, Grid<float>&, int64_t, int64_t, int, IndexWindow*, int*);
template bool build_index_window<double>(wtp_ctx*, IndexBuffers&, const double*, int64_t, int, Grid<double>&, int64_t, int64_t, int, IndexWindow*, int*);

}  // namespace wtp
