// bvh.cu — spacing callables on the device (src/discretization/spacings.jl).
//
// The variable spacings (LogLike :67-72, BoundaryLayerSpacing :121-133) need the distance
// from an arbitrary position to the nearest point of the spacing's own, static boundary
// set (_min_distance :19-23, a KDTree 1-NN in the reference). Queries may lie far from
// that set (deep interior points), where a uniform grid would walk many empty rings, so
// the boundary set gets a Morton-sorted implicit BVH instead: 8-point leaves, complete
// binary tree in heap layout, boxes built bottom-up; one thread answers one query with a
// short explicit stack. Built once per repel / spacing_eval call on the device.
#include "kernels.cuh"
#include "knn_core.cuh"

namespace wtp {

#define LAUNCH_CHECK(ctx)                         \
    do {                                          \
        (ctx)->launches++;                        \
        WTP_CUDA_CHECK(cudaPeekAtLastError());    \
    } while (0)

constexpr int BVH_LEAF = 8;

template <class T, int D>
__global__ void __launch_bounds__(256) morton_key_kernel(const T* __restrict__ pts, int64_t n, T lox, T loy, T loz, T sx, T sy, T sz,
                                                         uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int Q = D == 3 ? 1023 : 65535;
    auto quant = [&](T v, T lo, T s) {
        int q = (int)((v - lo) * s);
        return (uint32_t)(q < 0 ? 0 : (q > Q ? Q : q));
    };
    uint32_t qx = quant(pts[i * D + 0], lox, sx), qy = quant(pts[i * D + 1], loy, sy);
    uint32_t key;
    if (D == 3) key = spread3(qx) | (spread3(qy) << 1) | (spread3(quant(pts[i * D + (D - 1)], loz, sz)) << 2);
    else key = spread2(qx) | (spread2(qy) << 1);
    keys[i] = key;
    vals[i] = (uint32_t)i;
}

template <class T, int D>
__global__ void __launch_bounds__(256) bvh_gather_kernel(const T* __restrict__ pts, const uint32_t* __restrict__ vals, int64_t n,
                                                         P4<T>* __restrict__ sorted) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t i = vals[j];
    P4<T> p;
    p.x = pts[(size_t)i * D + 0];
    p.y = pts[(size_t)i * D + 1];
    p.z = D == 3 ? pts[(size_t)i * D + (D - 1)] : (T)0;
    p.w = idx_bits((T)0, i);
    sorted[j] = p;
}

template <class T>
__global__ void __launch_bounds__(256) bvh_leaf_kernel(const P4<T>* __restrict__ sorted, int64_t n, int64_t leaf_pow2, Box<T>* __restrict__ boxes) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= leaf_pow2) return;
    Box<T> bx;
    for (int d = 0; d < 3; ++d) { bx.lo[d] = t_inf<T>(); bx.hi[d] = -t_inf<T>(); }
    const int64_t j0 = b * BVH_LEAF, j1 = j0 + BVH_LEAF < n ? j0 + BVH_LEAF : n;
    for (int64_t j = j0; j < j1; ++j) {
        const P4<T> p = sorted[j];
        bx.lo[0] = p.x < bx.lo[0] ? p.x : bx.lo[0]; bx.hi[0] = p.x > bx.hi[0] ? p.x : bx.hi[0];
        bx.lo[1] = p.y < bx.lo[1] ? p.y : bx.lo[1]; bx.hi[1] = p.y > bx.hi[1] ? p.y : bx.hi[1];
        bx.lo[2] = p.z < bx.lo[2] ? p.z : bx.lo[2]; bx.hi[2] = p.z > bx.hi[2] ? p.z : bx.hi[2];
    }
    boxes[leaf_pow2 + b] = bx;
}

template <class T>
__global__ void __launch_bounds__(256) bvh_level_kernel(int64_t first, int64_t count, Box<T>* __restrict__ boxes) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int64_t i = first + t;
    const Box<T> a = boxes[2 * i], b = boxes[2 * i + 1];
    Box<T> o;
    for (int d = 0; d < 3; ++d) { o.lo[d] = a.lo[d] < b.lo[d] ? a.lo[d] : b.lo[d]; o.hi[d] = a.hi[d] > b.hi[d] ? a.hi[d] : b.hi[d]; }
    boxes[i] = o;
}

// inner boxes of a complete binary tree in heap layout whose P leaves (boxes[P..2P)) are set
template <class T>
void bvh_build_levels(wtp_ctx* ctx, Box<T>* boxes, int64_t P) {
    for (int64_t L = P / 2; L >= 1; L /= 2) {
        bvh_level_kernel<T><<<(unsigned)((L + 255) / 256), 256, 0, ctx->stream>>>(L, L, boxes);
        LAUNCH_CHECK(ctx);
    }
}
template void bvh_build_levels<float>(wtp_ctx*, Box<float>*, int64_t);
template void bvh_build_levels<double>(wtp_ctx*, Box<double>*, int64_t);

template <class T>
void bvh_build(wtp_ctx* ctx, BvhBuffers& bv, const T* d_bnd, int64_t n, int D) {
    WTP_REQUIRE(n > 0 && d_bnd, WTP_ERR_BAD_ARG, "variable spacing needs a non-empty boundary point set");
    WTP_REQUIRE(n < (int64_t)0xfffffff0u, WTP_ERR_BAD_ARG, "boundary set too large");
    double lo[3], hi[3];
    compute_bbox<T>(ctx, bv.ib, d_bnd, n, D, lo, hi);
    const double Q = D == 3 ? 1024.0 : 65536.0;
    T s[3], l[3];
    for (int d = 0; d < 3; ++d) {
        double ext = d < D ? hi[d] - lo[d] : 0.0;
        l[d] = (T)lo[d];
        s[d] = (T)(ext > 0 ? Q / ext * (1.0 - 1e-6) : 0.0);
    }
    uint32_t* keys = bv.ib.keys_a.as<uint32_t>((size_t)n);
    uint32_t* vals = bv.ib.vals_a.as<uint32_t>((size_t)n);
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (D == 2) morton_key_kernel<T, 2><<<nb, 256, 0, ctx->stream>>>(d_bnd, n, l[0], l[1], l[2], s[0], s[1], s[2], keys, vals);
    else morton_key_kernel<T, 3><<<nb, 256, 0, ctx->stream>>>(d_bnd, n, l[0], l[1], l[2], s[0], s[1], s[2], keys, vals);
    LAUNCH_CHECK(ctx);
    radix_sort_pairs(ctx, bv.ib, n, D == 3 ? 30 : 32);
    P4<T>* sorted = bv.ib.sorted.as<P4<T>>((size_t)n);
    if (D == 2) bvh_gather_kernel<T, 2><<<nb, 256, 0, ctx->stream>>>(d_bnd, bv.ib.vals_a.get<uint32_t>(), n, sorted);
    else bvh_gather_kernel<T, 3><<<nb, 256, 0, ctx->stream>>>(d_bnd, bv.ib.vals_a.get<uint32_t>(), n, sorted);
    LAUNCH_CHECK(ctx);
    const int64_t nleaf = (n + BVH_LEAF - 1) / BVH_LEAF;
    int64_t P = 1;
    while (P < nleaf) P <<= 1;
    Box<T>* boxes = bv.boxes.as<Box<T>>((size_t)2 * P);
    bvh_leaf_kernel<T><<<(unsigned)((P + 255) / 256), 256, 0, ctx->stream>>>(sorted, n, P, boxes);
    LAUNCH_CHECK(ctx);
    bvh_build_levels<T>(ctx, boxes, P);
    bv.n = n;
    bv.leaf_pow2 = P;
}
template void bvh_build<float>(wtp_ctx*, BvhBuffers&, const float*, int64_t, int);
template void bvh_build<double>(wtp_ctx*, BvhBuffers&, const double*, int64_t, int);

template <class T>
BvhView<T> bvh_view(const BvhBuffers& bv) {
    return BvhView<T>{bv.ib.sorted.get<P4<T>>(), bv.boxes.get<Box<T>>(), bv.n, bv.leaf_pow2};
}
template BvhView<float> bvh_view<float>(const BvhBuffers&);
template BvhView<double> bvh_view<double>(const BvhBuffers&);

// squared distance from q to the nearest boundary point; same d2 arithmetic as the k-NN
// `hint` (in/out): sorted position of a boundary point to start from (0xffffffff: none). A repel
// point moves little per iteration, so last iteration's nearest point gives a tight initial bound
// and the traversal prunes almost everything; the result is exact either way.
template <class T, int D>
__device__ __forceinline__ T bvh_nearest_d2(const BvhView<T>& bv, T qx, T qy, T qz, uint32_t& hint) {
    auto box_lb = [&](int64_t i) -> T {
        const Box<T> b = bv.boxes[i];
        T gx = qx < b.lo[0] ? sub_rn(b.lo[0], qx) : (qx > b.hi[0] ? sub_rn(qx, b.hi[0]) : (T)0);
        T gy = qy < b.lo[1] ? sub_rn(b.lo[1], qy) : (qy > b.hi[1] ? sub_rn(qy, b.hi[1]) : (T)0);
        T s = add_rn(mul_rn(gx, gx), mul_rn(gy, gy));
        if (D == 3) {
            T gz = qz < b.lo[2] ? sub_rn(b.lo[2], qz) : (qz > b.hi[2] ? sub_rn(qz, b.hi[2]) : (T)0);
            s = add_rn(s, mul_rn(gz, gz));
        }
        return s;  // +inf for empty nodes (lo = +inf)
    };
    T best = t_inf<T>();
    if (hint != 0xffffffffu) {
        const P4<T> p = load_p4<T>(bv.pts + hint);
        best = dist2_rn<T, D>(qx, qy, qz, p.x, p.y, p.z);
    }
    int stack[48];  // node ids < 2^31
    int sp = 0;
    stack[sp++] = 1;
    while (sp > 0) {
        const int64_t i = stack[--sp];
        if (i >= bv.leaf_pow2) {
            const int64_t j0 = (i - bv.leaf_pow2) * BVH_LEAF, j1 = j0 + BVH_LEAF < bv.n ? j0 + BVH_LEAF : bv.n;
            for (int64_t j = j0; j < j1; ++j) {
                const P4<T> p = load_p4<T>(bv.pts + j);
                const T d = dist2_rn<T, D>(qx, qy, qz, p.x, p.y, p.z);
                if (d < best) { best = d; hint = (uint32_t)j; }
            }
            continue;
        }
        const T ll = box_lb(2 * i), lr = box_lb(2 * i + 1);
        // push the farther child first so the nearer one is popped first
        if (ll <= lr) {
            if (lr < best) stack[sp++] = (int)(2 * i + 1);
            if (ll < best) stack[sp++] = (int)(2 * i);
        } else {
            if (ll < best) stack[sp++] = (int)(2 * i);
            if (lr < best) stack[sp++] = (int)(2 * i + 1);
        }
    }
    return best;
}

// The same search for the 32 queries of a warp at once (spacing_eval_ordered_kernel: the lanes are neighbours in space).
// One traversal per warp with warp-uniform control flow: a node is visited when ANY lane's bound reaches into its
// box, every lane keeps its own best; box and point loads are the same address on all lanes (one broadcast request
// instead of 32 diverging ones) and the stack is one per warp. Per lane the result is exact for the same reason as
// above: a subtree is skipped only when its box is no nearer than the lane's best at that moment, and bests only shrink.
// Idle lanes (live = false) never ask for a node. `wstack`: this warp's stack in shared memory (BVH_STACK entries).
constexpr int BVH_STACK = 64;
template <class T, int D>
__device__ __forceinline__ T bvh_nearest_d2_packet(const BvhView<T>& bv, T qx, T qy, T qz, bool live, uint32_t& hint, int* wstack) {
    const int lane = threadIdx.x & 31;
    auto box_lb = [&](int64_t i) -> T {
        const Box<T> b = bv.boxes[i];
        T gx = qx < b.lo[0] ? sub_rn(b.lo[0], qx) : (qx > b.hi[0] ? sub_rn(qx, b.hi[0]) : (T)0);
        T gy = qy < b.lo[1] ? sub_rn(b.lo[1], qy) : (qy > b.hi[1] ? sub_rn(qy, b.hi[1]) : (T)0);
        T s = add_rn(mul_rn(gx, gx), mul_rn(gy, gy));
        if (D == 3) {
            T gz = qz < b.lo[2] ? sub_rn(b.lo[2], qz) : (qz > b.hi[2] ? sub_rn(qz, b.hi[2]) : (T)0);
            s = add_rn(s, mul_rn(gz, gz));
        }
        return s;  // +inf for empty nodes (lo = +inf)
    };
    T best = live ? t_inf<T>() : (T)-1;                 // no box is nearer than -1: an idle lane wants nothing
    if (live && hint != 0xffffffffu) {
        const P4<T> p = load_p4<T>(bv.pts + hint);
        best = dist2_rn<T, D>(qx, qy, qz, p.x, p.y, p.z);
    }
    int sp = 0;
    int64_t node = 1;
    bool fresh = true;                                   // node comes straight from its parent's test (not from the stack)
    for (;;) {
        bool descend = true;
        if (!fresh) descend = __any_sync(0xffffffffu, box_lb(node) < best);   // the bests have shrunk since the push
        if (descend) {
            if (node >= bv.leaf_pow2) {
                const int64_t j0 = (node - bv.leaf_pow2) * BVH_LEAF, j1 = j0 + BVH_LEAF < bv.n ? j0 + BVH_LEAF : bv.n;
                for (int64_t j = j0; j < j1; ++j) {
                    const P4<T> p = load_p4<T>(bv.pts + j);
                    const T d = dist2_rn<T, D>(qx, qy, qz, p.x, p.y, p.z);
                    if (d < best) { best = d; hint = (uint32_t)j; }
                }
            } else {
                const T ll = box_lb(2 * node), lr = box_lb(2 * node + 1);
                const bool wl = ll < best, wr = lr < best;
                const unsigned ml = __ballot_sync(0xffffffffu, wl), mr = __ballot_sync(0xffffffffu, wr);
                if (ml != 0u && mr != 0u) {
                    // both children: first the one that is the nearer for most of the interested lanes
                    const unsigned left_nearer = __ballot_sync(0xffffffffu, wl && (!wr || ll <= lr));
                    const bool left_first = 2 * __popc(left_nearer) >= __popc(ml | mr);
                    if (lane == 0) wstack[sp] = (int)(2 * node + (left_first ? 1 : 0));
                    ++sp;
                    __syncwarp();
                    node = 2 * node + (left_first ? 0 : 1);
                    fresh = true;
                    continue;
                }
                if (ml != 0u || mr != 0u) { node = 2 * node + (ml != 0u ? 0 : 1); fresh = true; continue; }
            }
        }
        if (sp == 0) break;
        __syncwarp();
        node = wstack[--sp];
        fresh = false;
    }
    return best;
}

template <class T>
__device__ __forceinline__ T spacing_from_dmin(const SpacingP<T>& sp, T dmin) {
    if (sp.kind == WTP_SPACING_LOGLIKE) {                 // spacings.jl:67-72
        const T inv_growth = (T)1 - (sp.b - (T)1);
        const T aa = sp.a * inv_growth;
        return sp.a * dmin / (aa + dmin);
    }
    const T delta = sp.c;                                  // spacings.jl:121-133
    const T center = delta / (T)2, width = delta / (T)6;
    const T sigma = (T)1 / ((T)1 + exp(-(dmin - center) / width));
    return sp.a + (sp.b - sp.a) * sigma;
}

template <class T, int D>
__global__ void __launch_bounds__(128) spacing_eval_kernel(const SpacingP<T> sp, const BvhView<T> bv, const T* __restrict__ pts,
                                                           int64_t n, T* __restrict__ out, uint32_t* __restrict__ cache, int use_cache) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (sp.kind == WTP_SPACING_CONSTANT) { out[i] = sp.a; return; }
    const T qx = pts[i * D + 0], qy = pts[i * D + 1], qz = D == 3 ? pts[i * D + (D - 1)] : (T)0;
    uint32_t hint = (cache && use_cache) ? cache[i] : 0xffffffffu;
    const T dmin = sqrt(bvh_nearest_d2<T, D>(bv, qx, qy, qz, hint));
    if (cache) cache[i] = hint;
    out[i] = spacing_from_dmin<T>(sp, dmin);
}

// Same evaluation, one thread per record of a (slightly stale) spatially sorted copy of the snapshot:
// thread t takes the point whose caller index is stored in order[t] (movable points only), so that the
// lanes of a warp are neighbours in space and walk the BVH together instead of diverging.
template <class T, int D>
__global__ void __launch_bounds__(128) spacing_eval_ordered_kernel(const SpacingP<T> sp, const BvhView<T> bv, const T* __restrict__ pts,
                                                                   const P4<T>* __restrict__ order, int64_t n_order, uint32_t n_fixed, int64_t n,
                                                                   T* __restrict__ out, uint32_t* __restrict__ cache, int use_cache) {
    __shared__ int s_stack[128 / 32][BVH_STACK];
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // every lane stays for the warp's traversal; lanes without a movable point are idle
    int64_t i = -1;
    if (t < n_order) {
        const uint32_t g = idx_of(order[t]);
        if (g >= n_fixed && (int64_t)g - n_fixed < n) i = (int64_t)g - n_fixed;
    }
    const bool live = i >= 0;
    T qx = (T)0, qy = (T)0, qz = (T)0;
    uint32_t hint = 0xffffffffu;
    if (live) {
        qx = pts[i * D + 0]; qy = pts[i * D + 1]; qz = D == 3 ? pts[i * D + (D - 1)] : (T)0;
        if (cache && use_cache) hint = cache[i];
    }
    const T dmin = sqrt(bvh_nearest_d2_packet<T, D>(bv, qx, qy, qz, live, hint, s_stack[threadIdx.x >> 5]));
    if (!live) return;
    if (cache) cache[i] = hint;
    out[i] = spacing_from_dmin<T>(sp, dmin);
}

template <class T>
void spacing_eval_ordered(wtp_ctx* ctx, const SpacingP<T>& sp, const BvhBuffers& bv, const T* d_pts, int64_t n, int D, T* d_out,
                          uint32_t* d_nn_cache, bool use_cache, const P4<T>* d_order, int64_t n_order, int64_t n_fixed) {
    if (n <= 0 || n_order <= 0) return;
    const BvhView<T> v = bvh_view<T>(bv);
    const unsigned nb = (unsigned)((n_order + 127) / 128);
    if (D == 2) spacing_eval_ordered_kernel<T, 2><<<nb, 128, 0, ctx->stream>>>(sp, v, d_pts, d_order, n_order, (uint32_t)n_fixed, n, d_out, d_nn_cache, use_cache ? 1 : 0);
    else spacing_eval_ordered_kernel<T, 3><<<nb, 128, 0, ctx->stream>>>(sp, v, d_pts, d_order, n_order, (uint32_t)n_fixed, n, d_out, d_nn_cache, use_cache ? 1 : 0);
    LAUNCH_CHECK(ctx);
}
template void spacing_eval_ordered<float>(wtp_ctx*, const SpacingP<float>&, const BvhBuffers&, const float*, int64_t, int, float*, uint32_t*, bool, const P4<float>*, int64_t, int64_t);
template void spacing_eval_ordered<double>(wtp_ctx*, const SpacingP<double>&, const BvhBuffers&, const double*, int64_t, int, double*, uint32_t*, bool, const P4<double>*, int64_t, int64_t);

template <class T>
void spacing_eval(wtp_ctx* ctx, const SpacingP<T>& sp, const BvhBuffers& bv, const T* d_pts, int64_t n, int D, T* d_out,
                  uint32_t* d_nn_cache, bool use_cache) {
    if (n <= 0) return;
    WTP_REQUIRE(sp.kind == WTP_SPACING_CONSTANT || sp.kind == WTP_SPACING_LOGLIKE || sp.kind == WTP_SPACING_BOUNDARY_LAYER,
                WTP_ERR_UNSUPPORTED, "only ConstantSpacing, LogLike and BoundaryLayerSpacing run on the device");
    const BvhView<T> v = bvh_view<T>(bv);
    const unsigned nb = (unsigned)((n + 127) / 128);
    if (D == 2) spacing_eval_kernel<T, 2><<<nb, 128, 0, ctx->stream>>>(sp, v, d_pts, n, d_out, d_nn_cache, use_cache ? 1 : 0);
    else spacing_eval_kernel<T, 3><<<nb, 128, 0, ctx->stream>>>(sp, v, d_pts, n, d_out, d_nn_cache, use_cache ? 1 : 0);
    LAUNCH_CHECK(ctx);
}
template void spacing_eval<float>(wtp_ctx*, const SpacingP<float>&, const BvhBuffers&, const float*, int64_t, int, float*, uint32_t*, bool);
template void spacing_eval<double>(wtp_ctx*, const SpacingP<double>&, const BvhBuffers&, const double*, int64_t, int, double*, uint32_t*, bool);

// ------------------------------------------------------------------ forces
// compute_force(model, u) elementwise (src/repel_forces.jl:37, 57-60, 96-100, 124-127), the operations of the
// reference in its order, not contracted (this translation unit is compiled with -fmad=false). The repel sweeps
// (repel.cu) evaluate the same laws from u^2 with fused arithmetic.
template <class T>
__device__ __forceinline__ T force_fn(const ForceP<T> f, T u) {
    const T u2 = u * u;
    switch (f.kind) {
        case WTP_FORCE_INVERSE: { const T t = u2 + f.beta; return (T)1 / (t * t); }
        case WTP_FORCE_EQUILIBRIUM: { const T t = u2 + f.beta; return ((T)1 - u2) / (t * t); }
        case WTP_FORCE_CLIPPED: { const T t = u2 + f.beta; const T F = (f.u0 * f.u0 - u2) / (t * t); return F > (T)0 ? F : (T)0; }
        default: return ((T)1 - u2) / pow(u2 + f.beta, f.gamma);
    }
}
template <class T>
__global__ void __launch_bounds__(256) force_eval_kernel(const ForceP<T> f, const T* __restrict__ u, int64_t n, T* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = force_fn<T>(f, u[i]);
}
template <class T>
void force_eval(wtp_ctx* ctx, const ForceP<T>& f, const T* d_u, int64_t n, T* d_out) {
    if (n <= 0) return;
    force_eval_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(f, d_u, n, d_out);
    LAUNCH_CHECK(ctx);
}
template void force_eval<float>(wtp_ctx*, const ForceP<float>&, const float*, int64_t, float*);
template void force_eval<double>(wtp_ctx*, const ForceP<double>&, const double*, int64_t, double*);

}  // namespace wtp
