// consumers.cu — the other consumers of the k-NN index, as device entry points:
//
//   compute_normals(points; k)            src/normals.jl:9-44, 65-70   PCA normals (Hoppe 1992): per point the
//       eigenvector of the smallest eigenvalue of the covariance of its k nearest points (itself included)
//   _gradient_limit_field(...; k, tol, max_sweeps)   src/discretization/algorithms/octree.jl:677-717   the g-Lipschitz
//       envelope of a per-point field by min-plus Jacobi sweeps over the k-NN graph of the points
//
// Both are one k-NN pass of the index (knn.cu) followed by a thread-per-point kernel over the resulting table.
// orient_normals! (src/normals.jl:75-117) and split_surface! (src/surface_operations.jl:58-94) consume the same table
// (search(cloud, KNearestSearch), wtp_knn_self_*) and then walk a graph serially on the host; that part stays where it is.
//
// This translation unit is compiled with -fmad=false: h[j] + g * d is a multiply and an add, as in the reference.
#include <cmath>
#include <vector>

#include "kernels.cuh"
#include "knn_core.cuh"
#include "multi.h"

using namespace wtp;

namespace wtp {
int32_t fail(wtp_ctx* ctx, const Error& e);
template <class T>
void knn_device_self(wtp_ctx* ctx, const T* d_pts, int64_t N, int D, int k, int64_t* d_out_idx, T* d_out_dist);   // api.cu
}  // namespace wtp

#define LAUNCH_CHECK(ctx)                         \
    do {                                          \
        (ctx)->launches++;                        \
        WTP_CUDA_CHECK(cudaPeekAtLastError());    \
    } while (0)

#define API_BEGIN(ctx)                                                       \
    if (!(ctx)) return WTP_ERR_BAD_ARG;                                      \
    try {                                                                    \
        WTP_CUDA_CHECK(cudaSetDevice((ctx)->device));
#define API_END(ctx)                                                         \
    }                                                                        \
    catch (const Error& e) { return fail((ctx), e); }                        \
    catch (const std::bad_alloc&) { return fail((ctx), Error{WTP_ERR_OOM, "host allocation failed"}); } \
    catch (...) { return fail((ctx), Error{WTP_ERR_CUDA, "unknown failure"}); }  \
    return WTP_OK;

namespace wtp {

// ------------------------------------------------------------------ normals
// cov(v) of the k gathered points in T (mean first, then the centred second moments over k - 1: Statistics.cov), then
// the symmetric eigenproblem by cyclic Jacobi rotations in double: eigen(Symmetric(C)) sorts ascending, Q[:, 1] is the
// eigenvector of the smallest eigenvalue. Its sign is whatever LAPACK returns in the reference (compute_normals makes no
// promise, orient_normals! fixes it afterwards); here the first nonzero component is made positive.
template <int D>
__device__ __forceinline__ void smallest_eigenvector(double C[3][3], double* out) {
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0;
        for (int p = 0; p < D; ++p) for (int q = p + 1; q < D; ++q) off += C[p][q] * C[p][q];
        if (off == 0.0) break;
        for (int p = 0; p < D; ++p)
            for (int q = p + 1; q < D; ++q) {
                if (C[p][q] == 0.0) continue;
                const double theta = (C[q][q] - C[p][p]) / (2.0 * C[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int r = 0; r < D; ++r) {   // C <- C J
                    const double a = C[r][p], b = C[r][q];
                    C[r][p] = c * a - s * b; C[r][q] = s * a + c * b;
                }
                for (int r = 0; r < D; ++r) {   // C <- J^T C
                    const double a = C[p][r], b = C[q][r];
                    C[p][r] = c * a - s * b; C[q][r] = s * a + c * b;
                }
                for (int r = 0; r < D; ++r) {
                    const double a = V[r][p], b = V[r][q];
                    V[r][p] = c * a - s * b; V[r][q] = s * a + c * b;
                }
            }
    }
    int m = 0;
    for (int d = 1; d < D; ++d) if (C[d][d] < C[m][m]) m = d;
    double n2 = 0;
    for (int d = 0; d < D; ++d) n2 += V[d][m] * V[d][m];
    const double inv = 1.0 / sqrt(n2);
    double sign = 1.0;
    for (int d = 0; d < D; ++d) if (V[d][m] != 0.0) { sign = V[d][m] > 0 ? 1.0 : -1.0; break; }
    for (int d = 0; d < D; ++d) out[d] = sign * V[d][m] * inv;
}

template <class T, int D>
__global__ void __launch_bounds__(128) normals_kernel(const T* __restrict__ pts, const int64_t* __restrict__ idx, int64_t N, int k, T* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int64_t* row = idx + i * k;
    T mean[3] = {(T)0, (T)0, (T)0};
    for (int j = 0; j < k; ++j) {
        const int64_t n = row[j] - 1;
        for (int d = 0; d < D; ++d) mean[d] = mean[d] + pts[n * D + d];
    }
    for (int d = 0; d < D; ++d) mean[d] = mean[d] / (T)k;
    T S[3][3] = {};
    for (int j = 0; j < k; ++j) {
        const int64_t n = row[j] - 1;
        T c[3];
        for (int d = 0; d < D; ++d) c[d] = pts[n * D + d] - mean[d];
        for (int a = 0; a < D; ++a) for (int b = a; b < D; ++b) S[a][b] = S[a][b] + c[a] * c[b];
    }
    double C[3][3] = {};
    const T denom = (T)(k > 1 ? k - 1 : 1);
    for (int a = 0; a < D; ++a) for (int b = a; b < D; ++b) { C[a][b] = (double)(S[a][b] / denom); C[b][a] = C[a][b]; }
    double v[3];
    smallest_eigenvector<D>(C, v);
    for (int d = 0; d < D; ++d) out[i * D + d] = (T)v[d];
}

template <class T>
static int32_t normals_host(wtp_ctx* ctx, const T* pts, int64_t N, int32_t D, int32_t k, T* out) {
    API_BEGIN(ctx)
    WTP_REQUIRE(pts && out && N > 0, WTP_ERR_BAD_ARG, "null pointer or empty point set");
    WTP_REQUIRE(D == 2 || D == 3, WTP_ERR_BAD_ARG, "D must be 2 or 3");
    WTP_REQUIRE(k >= 1, WTP_ERR_BAD_ARG, "k must be >= 1");
    WTP_REQUIRE(ctx->world == 1, WTP_ERR_UNSUPPORTED, "compute_normals runs on a single-GPU context");
    if ((int64_t)k > N) k = (int32_t)N;                                                     // src/normals.jl:16
    T* d_pts = ctx->d_pts.as<T>((size_t)N * D);
    int64_t* d_idx = ctx->d_out_idx.as<int64_t>((size_t)N * k);
    T* d_out = ctx->d_out_dist.as<T>((size_t)N * D);
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_pts, pts, (size_t)N * D * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    knn_device_self<T>(ctx, d_pts, N, D, k, d_idx, nullptr);                                // search.(points, Ref(method)), :41
    const unsigned nb = (unsigned)((N + 127) / 128);
    if (D == 2) normals_kernel<T, 2><<<nb, 128, 0, ctx->stream>>>(d_pts, d_idx, N, k, d_out);
    else normals_kernel<T, 3><<<nb, 128, 0, ctx->stream>>>(d_pts, d_idx, N, k, d_out);
    LAUNCH_CHECK(ctx);
    WTP_CUDA_CHECK(cudaMemcpyAsync(out, d_out, (size_t)N * D * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END(ctx)
}

// -------------------------------------------------------- gradient-limit field
// One Jacobi sweep: hnew[a] = min(h[a], min_t h[nbr[t]] + g * d[t]) (octree.jl:694-702), and the largest relative
// change max_a |hnew[a] - h[a]| / h[a] (:704-707) as the bit pattern of a non-negative T through an integer atomic max.
template <class T>
__global__ void __launch_bounds__(256) gradient_sweep_kernel(const T* __restrict__ h, const int64_t* __restrict__ idx, const T* __restrict__ dist, int64_t n,
                                                             int k, T g, T* __restrict__ hnew, unsigned long long* __restrict__ maxrel_bits) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    T rel = (T)0;
    if (a < n) {
        const T h0 = h[a];
        T hi = h0;
        for (int t = 0; t < k; ++t) {
            const T cand = add_rn(h[idx[a * k + t] - 1], mul_rn(g, dist[a * k + t]));
            if (cand < hi) hi = cand;
        }
        hnew[a] = hi;
        rel = fabs(sub_rn(hi, h0)) / h0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const T r = __shfl_xor_sync(FULL, rel, o); rel = r > rel ? r : rel; }
    if ((threadIdx.x & 31) == 0 && rel > (T)0) {
        const unsigned long long bits = sizeof(T) == 4 ? (unsigned long long)__float_as_uint((float)rel) : (unsigned long long)__double_as_longlong((double)rel);
        atomicMax(maxrel_bits, bits);
    }
}

template <class T>
static int32_t gradient_limit_host(wtp_ctx* ctx, const T* centers, int64_t n, int32_t D, const T* h0, T g, int32_t k, double tol, int32_t max_sweeps,
                                   T* out, int32_t* sweeps_out) {
    API_BEGIN(ctx)
    WTP_REQUIRE(centers && h0 && out && n > 0, WTP_ERR_BAD_ARG, "null pointer or empty point set");
    WTP_REQUIRE(D == 2 || D == 3, WTP_ERR_BAD_ARG, "D must be 2 or 3");
    WTP_REQUIRE(k >= 1 && max_sweeps >= 0, WTP_ERR_BAD_ARG, "k must be >= 1 and max_sweeps >= 0");
    WTP_REQUIRE(ctx->world == 1, WTP_ERR_UNSUPPORTED, "the gradient-limit field runs on a single-GPU context");
    const int kk = (int)std::min<int64_t>(k, n);                                            // :684
    T* d_pts = ctx->d_pts.as<T>((size_t)n * D);
    int64_t* d_idx = ctx->d_out_idx.as<int64_t>((size_t)n * kk);
    T* d_dist = ctx->d_out_dist.as<T>((size_t)n * kk);
    T* d_h = ctx->d_misc.as<T>((size_t)2 * n);
    T* d_hn = d_h + n;
    unsigned long long* d_rel = ctx->d_reduce.as<unsigned long long>(1);
    unsigned long long* h_rel = static_cast<unsigned long long*>(ctx->h_pinned);
    cudaStream_t st = ctx->stream;
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_pts, centers, (size_t)n * D * sizeof(T), cudaMemcpyHostToDevice, st));
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_h, h0, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, st));
    knn_device_self<T>(ctx, d_pts, n, D, kk, d_idx, d_dist);                                // knn(tree, centers, kk, true), :686
    int sweeps = 0;
    const unsigned nb = (unsigned)((n + 255) / 256);
    for (int s = 0; s < max_sweeps; ++s) {                                                  // :693
        WTP_CUDA_CHECK(cudaMemsetAsync(d_rel, 0, sizeof(unsigned long long), st));
        gradient_sweep_kernel<T><<<nb, 256, 0, st>>>(d_h, d_idx, d_dist, n, kk, g, d_hn, d_rel);
        LAUNCH_CHECK(ctx);
        WTP_CUDA_CHECK(cudaMemcpyAsync(h_rel, d_rel, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        WTP_CUDA_CHECK(cudaStreamSynchronize(st));
        std::swap(d_h, d_hn);                                                               // copyto!(h, hnew), :708
        ++sweeps;
        double maxrel;
        if (sizeof(T) == 4) { const uint32_t b = (uint32_t)*h_rel; float f; memcpy(&f, &b, 4); maxrel = (double)f; }
        else { memcpy(&maxrel, h_rel, 8); }
        if (sizeof(T) == 4 ? (float)maxrel < (float)tol : maxrel < tol) break;             // :709 (maxrel is a T, tol a Real)
    }
    WTP_CUDA_CHECK(cudaMemcpyAsync(out, d_h, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, st));
    WTP_CUDA_CHECK(cudaStreamSynchronize(st));
    if (sweeps_out) *sweeps_out = sweeps;
    API_END(ctx)
}

}  // namespace wtp

extern "C" {

int32_t wtp_normals_f32(wtp_ctx* c, const float* p, int64_t N, int32_t D, int32_t k, float* out) { return normals_host<float>(solo_of(c), p, N, D, k, out); }
int32_t wtp_normals_f64(wtp_ctx* c, const double* p, int64_t N, int32_t D, int32_t k, double* out) { return normals_host<double>(solo_of(c), p, N, D, k, out); }
int32_t wtp_gradient_limit_f32(wtp_ctx* c, const float* centers, int64_t n, int32_t D, const float* h0, float g, int32_t k, double tol, int32_t max_sweeps,
                               float* out, int32_t* sweeps) {
    return gradient_limit_host<float>(solo_of(c), centers, n, D, h0, g, k, tol, max_sweeps, out, sweeps);
}
int32_t wtp_gradient_limit_f64(wtp_ctx* c, const double* centers, int64_t n, int32_t D, const double* h0, double g, int32_t k, double tol, int32_t max_sweeps,
                               double* out, int32_t* sweeps) {
    return gradient_limit_host<double>(solo_of(c), centers, n, D, h0, g, k, tol, max_sweeps, out, sweeps);
}

}  // extern "C"
