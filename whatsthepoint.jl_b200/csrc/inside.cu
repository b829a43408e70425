// inside.cu — point-in-domain tests against the boundary POINT cloud (not the mesh):
//   3-D  isinside(p, cloud) by the Green's-function sum over the boundary elements
//        g = sum_j area_j * (p - x_j) . n_j / |p - x_j|^3 ;  inside iff g < -2*pi      src/isinside.jl:86-106
//   2-D  isinside(p, polygon points) by the winding number: sum of the signed angles
//        subtended by consecutive boundary points; inside iff |sum| >= 1e3*eps, points
//        coincident with a boundary point (distance < 1e2*eps) count as inside       src/isinside.jl:18-35
// These are the survivor filter of repel(cloud, spacing) (src/repel.jl:90): O(n_vol * n_bnd), dense and
// embarrassingly parallel. One thread per test point; the boundary elements stream through shared memory
// in tiles that the whole CTA loads together.
#include <cmath>

#include "kernels.cuh"

namespace wtp {

#define LAUNCH_CHECK(ctx)                         \
    do {                                          \
        (ctx)->launches++;                        \
        WTP_CUDA_CHECK(cudaPeekAtLastError());    \
    } while (0)

constexpr int IN_THREADS = 256;
constexpr int IN_TILE = 256;

// Sum order: each tile is added up element by element into a tile partial, tile partials are added in order.
template <class T>
__global__ void __launch_bounds__(IN_THREADS) greens_kernel(const T* __restrict__ pts, int64_t n, const T* __restrict__ bx, const T* __restrict__ bn,
                                                            const T* __restrict__ ba, int64_t m, T* __restrict__ g_out, uint8_t* __restrict__ inside) {
    __shared__ T s[IN_TILE][7];
    const int64_t i = (int64_t)blockIdx.x * IN_THREADS + threadIdx.x;
    const bool live = i < n;
    const T px = live ? pts[i * 3] : (T)0, py = live ? pts[i * 3 + 1] : (T)0, pz = live ? pts[i * 3 + 2] : (T)0;
    T g = (T)0;
    for (int64_t j0 = 0; j0 < m; j0 += IN_TILE) {
        const int64_t j = j0 + threadIdx.x;
        __syncthreads();
        if (j < m) {
            s[threadIdx.x][0] = bx[j * 3]; s[threadIdx.x][1] = bx[j * 3 + 1]; s[threadIdx.x][2] = bx[j * 3 + 2];
            s[threadIdx.x][3] = bn[j * 3]; s[threadIdx.x][4] = bn[j * 3 + 1]; s[threadIdx.x][5] = bn[j * 3 + 2];
            s[threadIdx.x][6] = ba[j];
        }
        __syncthreads();
        const int cnt = (int)(m - j0 < IN_TILE ? m - j0 : IN_TILE);
        T part = (T)0;
#pragma unroll 4
        for (int t = 0; t < cnt; ++t) {
            const T dx = px - s[t][0], dy = py - s[t][1], dz = pz - s[t][2];
            const T dn = sqrt((dx * dx + dy * dy) + dz * dz);
            const T dot = (dx * s[t][3] + dy * s[t][4]) + dz * s[t][5];
            part = part + s[t][6] * dot / (dn * dn * dn);                      // area * dist . normal / norm(dist)^3
        }
        g = g + part;
    }
    if (live) {
        if (g_out) g_out[i] = g;
        inside[i] = g < (T)(-6.283185307179586) ? 1 : 0;                      // the -4*pi of the Green's function is in the inequality
    }
}

template <class T>
__global__ void __launch_bounds__(IN_THREADS) winding_kernel(const T* __restrict__ pts, int64_t n, const T* __restrict__ poly, int64_t m,
                                                             T* __restrict__ w_out, uint8_t* __restrict__ inside) {
    __shared__ T s[IN_TILE + 1][2];
    const int64_t i = (int64_t)blockIdx.x * IN_THREADS + threadIdx.x;
    const bool live = i < n;
    const T px = live ? pts[i * 2] : (T)0, py = live ? pts[i * 2 + 1] : (T)0;
    const T eps = sizeof(T) == 4 ? (T)1.1920929e-7 : (T)2.220446049250313e-16;
    T sum = (T)0;
    bool on_boundary = false;
    for (int64_t j0 = 0; j0 < m; j0 += IN_TILE) {
        __syncthreads();
        for (int t = threadIdx.x; t <= IN_TILE; t += IN_THREADS) {           // the tile plus the next vertex (wrapping to the first)
            const int64_t j = (j0 + t) % m;
            if (j0 + t <= m) { s[t][0] = poly[j * 2]; s[t][1] = poly[j * 2 + 1]; }
        }
        __syncthreads();
        const int cnt = (int)(m - j0 < IN_TILE ? m - j0 : IN_TILE);
        T part = (T)0;
        for (int t = 0; t < cnt; ++t) {
            const T ux = s[t][0] - px, uy = s[t][1] - py, vx = s[t + 1][0] - px, vy = s[t + 1][1] - py;
            on_boundary = on_boundary || sqrt(ux * ux + uy * uy) < (T)1.0e2 * eps;
            part = part + atan2(ux * vy - uy * vx, ux * vx + uy * vy);         // signed angle A-p-B
        }
        sum = sum + part;
    }
    if (live) {
        if (w_out) w_out[i] = sum;
        inside[i] = (on_boundary || !(fabs(sum) < (T)1.0e3 * eps)) ? 1 : 0;
    }
}

template <class T>
void greens_isinside(wtp_ctx* ctx, const T* d_pts, int64_t n, const T* d_bx, const T* d_bn, const T* d_ba, int64_t m, T* d_g, uint8_t* d_inside) {
    if (n <= 0) return;
    greens_kernel<T><<<(unsigned)((n + IN_THREADS - 1) / IN_THREADS), IN_THREADS, 0, ctx->stream>>>(d_pts, n, d_bx, d_bn, d_ba, m, d_g, d_inside);
    LAUNCH_CHECK(ctx);
}
template void greens_isinside<float>(wtp_ctx*, const float*, int64_t, const float*, const float*, const float*, int64_t, float*, uint8_t*);
template void greens_isinside<double>(wtp_ctx*, const double*, int64_t, const double*, const double*, const double*, int64_t, double*, uint8_t*);

template <class T>
void winding_isinside(wtp_ctx* ctx, const T* d_pts, int64_t n, const T* d_poly, int64_t m, T* d_w, uint8_t* d_inside) {
    if (n <= 0) return;
    winding_kernel<T><<<(unsigned)((n + IN_THREADS - 1) / IN_THREADS), IN_THREADS, 0, ctx->stream>>>(d_pts, n, d_poly, m, d_w, d_inside);
    LAUNCH_CHECK(ctx);
}
template void winding_isinside<float>(wtp_ctx*, const float*, int64_t, const float*, int64_t, float*, uint8_t*);
template void winding_isinside<double>(wtp_ctx*, const double*, int64_t, const double*, int64_t, double*, uint8_t*);

}  // namespace wtp
