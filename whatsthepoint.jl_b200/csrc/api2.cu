// api2.cu — C ABI entry points for radius topology, repel, spacing / force evaluation
// and cloud metrics (include/wtp_cuda.h).
#include <cmath>
#include <cstdlib>
#include <new>

#include "kernels.cuh"
#include "knn_core.cuh"
#include "multi.h"

using namespace wtp;

namespace wtp {
int32_t fail(wtp_ctx* ctx, const Error& e);
void finish_timing(wtp_ctx* ctx, int sort_passes, int query_launches, int64_t n_cells, int64_t n_expanded);
template <class T>
void relax_device(wtp_ctx* ctx, T* d_snap, int64_t n_fixed, int64_t n_move, int D, const wtp_spacing* sp_in, const T* d_bnd,
                  const wtp_force* fm, const wtp_repel_params* prm, MeshBuffers* mesh, T* conv, wtp_trace_entry* trace,
                  wtp_repel_result* res);
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

void d2h_widen_u32(wtp_ctx* ctx, const uint32_t* d_src, size_t n, int64_t* h_dst, uint64_t max_value);   // api.cu

__global__ void __launch_bounds__(256) narrow_indices_kernel(const int64_t* __restrict__ in, size_t n, uint32_t* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (uint32_t)in[i];
}
static void narrow_indices(wtp_ctx* ctx, const int64_t* d_in, size_t n, uint32_t* d_out) {
    const unsigned nb = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)kNumSMs * 16);
    narrow_indices_kernel<<<nb, 256, 0, ctx->stream>>>(d_in, n, d_out);
    LAUNCH_CHECK(ctx);
}

// ----------------------------------------------------------------- radius
template <class T>
static void radius_count_device(wtp_ctx* ctx, const T* d_pts, int64_t N, int D, T r, int64_t* d_offsets) {
    WTP_REQUIRE(D == 2 || D == 3, WTP_ERR_BAD_ARG, "D must be 2 or 3");
    WTP_REQUIRE(r >= (T)0 && std::isfinite((double)r), WTP_ERR_BAD_ARG, "radius must be finite and >= 0");
    IndexBuffers& ib = ctx->index[0];
    double lo[3], hi[3];
    compute_bbox<T>(ctx, ib, d_pts, N, D, lo, hi);
    // cell size >= r (with margin) so the 3^D block around a query contains every hit; as close to r as possible
    // (fewer candidates per hit), never below two points per cell on average (tiny radii)
    Grid<T> g = make_grid<T>(N, D, lo, hi, ctx->cell_occupancy > 0 ? ctx->cell_occupancy : 2.0, (double)r * 1.001, 0);
    int passes = build_index<T>(ctx, ib, d_pts, N, D, g);
    const int64_t qb = wtp_shard_begin(N, ctx->rank, ctx->world), qe = wtp_shard_end(N, ctx->rank, ctx->world);
    // (sharded: the tiled passes take the caller range as a keep filter; only the general kernels on their own need a list)
    const uint32_t* qlist = nullptr;
    if (ctx->world > 1 && std::getenv("WTP_NO_TILED") != nullptr) {
        build_query_list(ctx, ib, N, qb, qe, sizeof(T) == 8, ctx->d_misc, ctx->d_misc2, ctx->d_qlist);
        qlist = ctx->d_qlist.get<uint32_t>();
    }
    const int64_t nq = qe - qb;
    uint32_t* counts = ctx->d_counts.as<uint32_t>((size_t)nq + 1);
    radius_count<T>(ctx, ib, g, N, D, r, qlist, nq, qb, counts);
    {
        ScopedPhase ph(ctx->timer, PH_SCAN);
        exclusive_scan_u32_to_i64(ctx, ib.scan_tmp, counts, d_offsets, nq);
    }
    int64_t* h_nnz = static_cast<int64_t*>(ctx->h_pinned);
    WTP_CUDA_CHECK(cudaMemcpyAsync(h_nnz, d_offsets + nq, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    auto& st = ctx->radius;
    st.pending = true; st.f64 = sizeof(T) == 8; st.N = N; st.D = D; st.r = (double)r; st.nnz = *h_nnz;
    st.q_begin = qb; st.q_end = qe;
    static_assert(sizeof(Grid<T>) <= sizeof(ctx->grid_storage[0]), "grid storage too small");
    memcpy(ctx->grid_storage[0], &g, sizeof(g));
    finish_timing(ctx, passes, 1, g.ncells, 0);
}

template <class T>
static void radius_fill_device(wtp_ctx* ctx, const int64_t* d_offsets, int64_t* d_indices) {
    auto& st = ctx->radius;
    Grid<T> g;
    memcpy(&g, ctx->grid_storage[0], sizeof(g));
    const uint32_t* qlist = ctx->world > 1 && std::getenv("WTP_NO_TILED") != nullptr ? ctx->d_qlist.get<uint32_t>() : nullptr;
    ctx->d_misc2.as<uint32_t>((size_t)std::max<int64_t>(st.nnz, 1));
    radius_fill<T>(ctx, ctx->index[0], g, st.N, st.D, (T)st.r, qlist, st.q_end - st.q_begin, st.q_begin, d_offsets, d_indices);
}

template <class T>
static int32_t radius_count_host(wtp_ctx* ctx, const T* pts, int64_t N, int32_t D, T r, int64_t* offsets) {
    API_BEGIN(ctx)
    ctx->radius.pending = false;
    WTP_REQUIRE(pts && offsets && N > 0, WTP_ERR_BAD_ARG, "null pointer or empty point set");
    WTP_REQUIRE(D == 2 || D == 3, WTP_ERR_BAD_ARG, "D must be 2 or 3");
    ctx->timer.reset(ctx->stream);
    ctx->timer.begin_total();
    T* d_pts = ctx->d_pts.as<T>((size_t)N * D);
    {
        ScopedPhase ph(ctx->timer, PH_H2D);
        WTP_CUDA_CHECK(cudaMemcpyAsync(d_pts, pts, (size_t)N * D * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    }
    const int64_t nq = wtp_shard_end(N, ctx->rank, ctx->world) - wtp_shard_begin(N, ctx->rank, ctx->world);
    int64_t* d_off = ctx->d_offsets.as<int64_t>((size_t)nq + 1);
    radius_count_device<T>(ctx, d_pts, N, D, r, d_off);
    wtp_timing keep = ctx->last_timing;
    {
        ScopedPhase ph(ctx->timer, PH_D2H);
        WTP_CUDA_CHECK(cudaMemcpyAsync(offsets, d_off, (size_t)(nq + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    ctx->timer.end_total();
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->last_timing = keep;
    ctx->radius.dev_input = false;
    API_END(ctx)
}

template <class T>
static int32_t radius_count_dev(wtp_ctx* ctx, const T* d_pts, int64_t N, int32_t D, T r, int64_t* d_offsets) {
    API_BEGIN(ctx)
    ctx->radius.pending = false;
    WTP_REQUIRE(d_pts && d_offsets && N > 0, WTP_ERR_BAD_ARG, "null pointer or empty point set");
    ctx->timer.reset(ctx->stream);
    ctx->timer.begin_total();
    radius_count_device<T>(ctx, d_pts, N, D, r, d_offsets);
    ctx->timer.end_total();
    ctx->radius.dev_input = true;
    ctx->radius.pts = d_offsets;
    API_END(ctx)
}

// The CSR of a multi-device context: every child counts the rows of its contiguous caller range into its own offsets
// (0-based within the shard); the parent shifts them into one global prefix and remembers where each child's entries
// start, so that wtp_radius_fill lets every child fill its part of the caller's indices array.
template <class T>
static int32_t radius_count_entry(wtp_ctx* c, const T* p, int64_t N, int32_t D, T r, int64_t* off) {
    if (!is_multi(c) || N < (int64_t)4096 * (int64_t)c->children.size()) {
        if (is_multi(c)) c->radius.pending = false;
        return radius_count_host<T>(solo_of(c), p, N, D, r, off);
    }
    c->radius.pending = false;
    if (!off) return fail(c, Error{WTP_ERR_BAD_ARG, "null offsets"});
    const int G = (int)c->children.size();
    std::vector<std::vector<int64_t>> local((size_t)G);
    for (int g = 0; g < G; ++g) local[(size_t)g].resize((size_t)(wtp_shard_end(N, g, G) - wtp_shard_begin(N, g, G) + 1));
    const int32_t rc = multi_run(c, [&](wtp_ctx* ch, int g) { return radius_count_host<T>(ch, p, N, D, r, local[(size_t)g].data()); });
    if (rc != 0) return rc;
    c->children_base.assign((size_t)G, 0);
    int64_t base = 0;
    for (int g = 0; g < G; ++g) {
        const int64_t qb = wtp_shard_begin(N, g, G), nq = wtp_shard_end(N, g, G) - qb;
        c->children_base[(size_t)g] = base;
        for (int64_t t = 0; t < nq; ++t) off[qb + t] = base + local[(size_t)g][(size_t)t];
        base += local[(size_t)g][(size_t)nq];
    }
    off[N] = base;
    c->radius.pending = true;
    c->radius.nnz = base;
    return WTP_OK;
}
static int32_t no_device_pointers2(wtp_ctx* c) {
    return fail(c, Error{WTP_ERR_UNSUPPORTED, "device-pointer entry points belong to one device: use a single-device context"});
}

}  // namespace wtp

extern "C" {

int32_t wtp_radius_count_f32(wtp_ctx* c, const float* p, int64_t N, int32_t D, float r, int64_t* off) { return radius_count_entry<float>(c, p, N, D, r, off); }
int32_t wtp_radius_count_f64(wtp_ctx* c, const double* p, int64_t N, int32_t D, double r, int64_t* off) { return radius_count_entry<double>(c, p, N, D, r, off); }
int32_t wtp_radius_count_dev_f32(wtp_ctx* c, const float* p, int64_t N, int32_t D, float r, int64_t* off) { return is_multi(c) ? no_device_pointers2(c) : radius_count_dev<float>(c, p, N, D, r, off); }
int32_t wtp_radius_count_dev_f64(wtp_ctx* c, const double* p, int64_t N, int32_t D, double r, int64_t* off) { return is_multi(c) ? no_device_pointers2(c) : radius_count_dev<double>(c, p, N, D, r, off); }

int64_t wtp_radius_nnz(const wtp_ctx* ctx) {
    if (is_multi(ctx)) {
        if (ctx->radius.pending) return ctx->radius.nnz;                 // the sharded count's total
        ctx = ctx->solo;
    }
    return ctx && ctx->radius.pending ? ctx->radius.nnz : -1;
}

int32_t wtp_radius_fill(wtp_ctx* ctx, int64_t* indices) {
    if (is_multi(ctx)) {
        if (!ctx->radius.pending) return wtp_radius_fill(ctx->solo, indices);
        ctx->radius.pending = false;
        // every child fills its part of the CSR at the global offset the count pass gave it (kept in q_begin of the parent's
        // state per child: see radius_count_entry)
        return multi_run(ctx, [&](wtp_ctx* ch, int r) { return wtp_radius_fill(ch, indices ? indices + ctx->children_base[(size_t)r] : nullptr); });
    }
    API_BEGIN(ctx)
    auto& st = ctx->radius;
    WTP_REQUIRE(st.pending && !st.dev_input, WTP_ERR_STATE, "wtp_radius_fill must directly follow wtp_radius_count_* on the same context");
    WTP_REQUIRE(indices || st.nnz == 0, WTP_ERR_BAD_ARG, "null indices");
    st.pending = false;
    if (st.nnz > 0) {
        ctx->timer.reset(ctx->stream);
        ctx->timer.begin_total();
        int64_t* d_ind = ctx->d_indices.as<int64_t>((size_t)st.nnz);
        if (st.f64) radius_fill_device<double>(ctx, ctx->d_offsets.get<int64_t>(), d_ind);
        else radius_fill_device<float>(ctx, ctx->d_offsets.get<int64_t>(), d_ind);
        {
            ScopedPhase ph(ctx->timer, PH_D2H);
            if (st.nnz < ((int64_t)4 << 20)) {
                WTP_CUDA_CHECK(cudaMemcpyAsync(indices, d_ind, (size_t)st.nnz * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
            } else {
                // large result: 4 bytes per entry cross PCIe, widened into the caller's array by the host pool (api.cu)
                uint32_t* d_ind32 = ctx->d_out_idx.as<uint32_t>((size_t)st.nnz);
                narrow_indices(ctx, d_ind, (size_t)st.nnz, d_ind32);
                d2h_widen_u32(ctx, d_ind32, (size_t)st.nnz, indices, (uint64_t)st.N);
            }
        }
        ctx->timer.end_total();
        WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    API_END(ctx)
}

int32_t wtp_radius_fill_dev(wtp_ctx* ctx, int64_t* d_indices) {
    if (is_multi(ctx)) return no_device_pointers2(ctx);
    API_BEGIN(ctx)
    auto& st = ctx->radius;
    WTP_REQUIRE(st.pending && st.dev_input, WTP_ERR_STATE, "wtp_radius_fill_dev must directly follow wtp_radius_count_dev_* on the same context");
    WTP_REQUIRE(d_indices || st.nnz == 0, WTP_ERR_BAD_ARG, "null indices");
    st.pending = false;
    if (st.nnz > 0) {
        ctx->timer.reset(ctx->stream);
        ctx->timer.begin_total();
        const int64_t* d_off = static_cast<const int64_t*>(st.pts);
        if (st.f64) radius_fill_device<double>(ctx, d_off, d_indices);
        else radius_fill_device<float>(ctx, d_off, d_indices);
        ctx->timer.end_total();
    }
    API_END(ctx)
}

}  // extern "C"

// ------------------------------------------------------------------ repel
namespace wtp {

template <class T>
static int32_t repel_host(wtp_ctx* ctx, T* snap, int64_t n_fixed, int64_t n_move, int32_t D, const wtp_spacing* sp,
                          const wtp_force* fm, const wtp_repel_params* prm, const wtp_wall_mesh* wall, T* conv,
                          wtp_trace_entry* trace, wtp_repel_result* res) {
    API_BEGIN(ctx)
    WTP_REQUIRE(snap && sp && fm && prm && conv && res, WTP_ERR_BAD_ARG, "null pointer");
    WTP_REQUIRE(D == 2 || D == 3, WTP_ERR_BAD_ARG, "D must be 2 or 3");
    WTP_REQUIRE(n_fixed >= 0 && n_move >= 0 && n_fixed + n_move > 0, WTP_ERR_BAD_ARG, "empty snapshot");
    WTP_REQUIRE((wall != nullptr) == (prm->wall == WTP_WALL_MESH), WTP_ERR_BAD_ARG, "params.wall and the wall mesh argument disagree");
    WTP_REQUIRE(!wall || D == 3, WTP_ERR_BAD_ARG, "the mesh wall rule is 3-D only (src/repel.jl:123)");
    WTP_REQUIRE(!wall || n_move == 0 || wall->is_bnd, WTP_ERR_BAD_ARG, "mesh wall needs the is_bnd flags");
    ctx->timer.reset(ctx->stream);
    ctx->timer.begin_total();
    const int64_t n_all = n_fixed + n_move;
    T* d_snap = ctx->d_pts.as<T>((size_t)n_all * D);
    const T* d_bnd = nullptr;
    {
        ScopedPhase ph(ctx->timer, PH_H2D);
        WTP_CUDA_CHECK(cudaMemcpyAsync(d_snap, snap, (size_t)n_all * D * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        if (sp->kind != WTP_SPACING_CONSTANT) {
            WTP_REQUIRE(sp->bnd_pts && sp->n_bnd > 0, WTP_ERR_BAD_ARG, "variable spacing needs its boundary point set");
            T* b = ctx->d_spacing_pts.as<T>((size_t)sp->n_bnd * D);
            WTP_CUDA_CHECK(cudaMemcpyAsync(b, sp->bnd_pts, (size_t)sp->n_bnd * D * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
            d_bnd = b;
        }
    }
    MeshBuffers* mesh = nullptr;
    if (wall) {   // TriangleIndex arrays -> device BVH; per-point wall state (src/repel.jl:151-156)
        mesh = &ctx->mesh;
        mesh_build<T>(ctx, *mesh, wall);
        const size_t nm = (size_t)std::max<int64_t>(n_move, 1);
        WTP_CUDA_CHECK(cudaMemcpyAsync(mesh->is_bnd.as<uint8_t>(nm), wall->is_bnd, (size_t)n_move, cudaMemcpyHostToDevice, ctx->stream));
        WTP_CUDA_CHECK(cudaMemsetAsync(mesh->tri_idx.as<int64_t>(nm), 0, nm * sizeof(int64_t), ctx->stream));
        WTP_CUDA_CHECK(cudaMemsetAsync(mesh->escaped.as<uint8_t>(nm), 0, nm, ctx->stream));
        WTP_CUDA_CHECK(cudaMemsetAsync(mesh->hint.as<uint32_t>(nm), 0xff, nm * sizeof(uint32_t), ctx->stream));
    }
    relax_device<T>(ctx, d_snap, n_fixed, n_move, D, sp, d_bnd, fm, prm, mesh, conv, trace, res);
    wtp_timing keep = ctx->last_timing;
    if (!ctx->quiet) {   // (a multi-device context: every child holds the same final state, rank 0 alone writes it back)
        ScopedPhase ph(ctx->timer, PH_D2H);
        WTP_CUDA_CHECK(cudaMemcpyAsync(snap + (size_t)n_fixed * D, d_snap + (size_t)n_fixed * D, (size_t)n_move * D * sizeof(T),
                                       cudaMemcpyDeviceToHost, ctx->stream));
        if (mesh && wall->tri_indices) WTP_CUDA_CHECK(cudaMemcpyAsync(wall->tri_indices, mesh->tri_idx.get<int64_t>(), (size_t)n_move * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (mesh && wall->escaped) WTP_CUDA_CHECK(cudaMemcpyAsync(wall->escaped, mesh->escaped.get<uint8_t>(), (size_t)n_move, cudaMemcpyDeviceToHost, ctx->stream));
        if (mesh && prm->deposit_ratio > 0)   // deposition converts volume points into boundary points (src/repel.jl:509)
            WTP_CUDA_CHECK(cudaMemcpyAsync(const_cast<uint8_t*>(wall->is_bnd), mesh->is_bnd.get<uint8_t>(), (size_t)n_move, cudaMemcpyDeviceToHost, ctx->stream));
    }
    ctx->timer.end_total();
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->last_timing = keep;
    API_END(ctx)
}

template <class T>
static int32_t repel_dev(wtp_ctx* ctx, T* d_snap, int64_t n_fixed, int64_t n_move, int32_t D, const wtp_spacing* sp,
                         const wtp_force* fm, const wtp_repel_params* prm, T* conv, wtp_trace_entry* trace, wtp_repel_result* res) {
    API_BEGIN(ctx)
    WTP_REQUIRE(d_snap && sp && fm && prm && conv && res, WTP_ERR_BAD_ARG, "null pointer");
    ctx->timer.reset(ctx->stream);
    ctx->timer.begin_total();
    relax_device<T>(ctx, d_snap, n_fixed, n_move, D, sp, static_cast<const T*>(sp->bnd_pts), fm, prm, nullptr, conv, trace, res);
    ctx->timer.end_total();
    API_END(ctx)
}

template <class T>
static int32_t spacing_eval_host(wtp_ctx* ctx, const wtp_spacing* sp, const T* pts, int64_t N, int32_t D, T* out) {
    API_BEGIN(ctx)
    WTP_REQUIRE(sp && pts && out && N > 0, WTP_ERR_BAD_ARG, "null pointer or empty point set");
    WTP_REQUIRE(D == 2 || D == 3, WTP_ERR_BAD_ARG, "D must be 2 or 3");
    WTP_REQUIRE(sp->kind >= WTP_SPACING_CONSTANT && sp->kind <= WTP_SPACING_BOUNDARY_LAYER, WTP_ERR_UNSUPPORTED, "user-defined spacing callable cannot cross the C ABI");
    T* d_pts = ctx->d_pts.as<T>((size_t)N * D);
    T* d_out = ctx->d_spacings.as<T>((size_t)N);
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_pts, pts, (size_t)N * D * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    if (sp->kind != WTP_SPACING_CONSTANT) {
        WTP_REQUIRE(sp->bnd_pts && sp->n_bnd > 0, WTP_ERR_BAD_ARG, "variable spacing needs its boundary point set");
        T* b = ctx->d_spacing_pts.as<T>((size_t)sp->n_bnd * D);
        WTP_CUDA_CHECK(cudaMemcpyAsync(b, sp->bnd_pts, (size_t)sp->n_bnd * D * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        bvh_build<T>(ctx, ctx->bvh, b, sp->n_bnd, D);
    }
    const SpacingP<T> spp{sp->kind, (T)sp->a, (T)sp->b, (T)sp->c};
    spacing_eval<T>(ctx, spp, ctx->bvh, d_pts, N, D, d_out);
    WTP_CUDA_CHECK(cudaMemcpyAsync(out, d_out, (size_t)N * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END(ctx)
}

template <class T>
static int32_t force_eval_host(wtp_ctx* ctx, const wtp_force* f, const T* u, int64_t n, T* out) {
    API_BEGIN(ctx)
    WTP_REQUIRE(f && u && out && n > 0, WTP_ERR_BAD_ARG, "null pointer or empty input");
    WTP_REQUIRE(f->kind >= WTP_FORCE_INVERSE && f->kind <= WTP_FORCE_STRONG, WTP_ERR_UNSUPPORTED, "user-defined RepelForceModel cannot cross the C ABI");
    T* d_u = ctx->d_misc.as<T>((size_t)n);
    T* d_o = ctx->d_misc2.as<T>((size_t)n);
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_u, u, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    const ForceP<T> fp{f->kind, (T)f->beta, (T)f->u0, (T)f->gamma};
    force_eval<T>(ctx, fp, d_u, n, d_o);
    WTP_CUDA_CHECK(cudaMemcpyAsync(out, d_o, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END(ctx)
}

// repel on a multi-device context: the children run the same call as the ranks of one communicator (sharded sweeps,
// exchanged runs, identical stop decisions); every child ends with the whole final state, rank 0 writes it back. Small
// problems, the mesh wall and deposition (whose host-side passes are per process) run on the single-device context.
template <class T>
static int32_t repel_entry(wtp_ctx* c, T* snap, int64_t nf, int64_t nm, int32_t D, const wtp_spacing* sp, const wtp_force* fm,
                           const wtp_repel_params* prm, const wtp_wall_mesh* wall, T* conv, wtp_trace_entry* tr, wtp_repel_result* res) {
    if (!is_multi(c) || wall || !prm || nf + nm < (int64_t)4096 * (int64_t)c->children.size())
        return repel_host<T>(solo_of(c), snap, nf, nm, D, sp, fm, prm, wall, conv, tr, res);
    const int G = (int)c->children.size();
    const size_t slots = (size_t)std::max(prm->max_iters, 1);
    std::vector<std::vector<T>> conv_x((size_t)G);
    std::vector<std::vector<wtp_trace_entry>> tr_x((size_t)G);
    std::vector<wtp_repel_result> res_x((size_t)G);
    for (int g = 1; g < G; ++g) { conv_x[(size_t)g].resize(slots); if (tr) tr_x[(size_t)g].resize(slots); }
    return multi_run(c, [&](wtp_ctx* ch, int g) {
        ch->quiet = g > 0;
        const int32_t rc = g == 0 ? repel_host<T>(ch, snap, nf, nm, D, sp, fm, prm, nullptr, conv, tr, res)
                                  : repel_host<T>(ch, snap, nf, nm, D, sp, fm, prm, nullptr, conv_x[(size_t)g].data(),
                                                  tr ? tr_x[(size_t)g].data() : nullptr, &res_x[(size_t)g]);
        ch->quiet = false;
        return rc;
    });
}

// ----------------------------------------------------------- mesh queries
// isinside(points, octree) / _project_to_boundary for a batch of host points (3-D)
template <class T>
static int32_t mesh_query_host(wtp_ctx* ctx, const wtp_wall_mesh* wall, const T* pts, int64_t N, uint8_t* out_inside, T* out_pts, int64_t* out_tri) {
    API_BEGIN(ctx)
    WTP_REQUIRE(wall && pts && N > 0 && (out_inside || (out_pts && out_tri)), WTP_ERR_BAD_ARG, "null pointer or empty point set");
    MeshBuffers& mb = ctx->mesh;
    mesh_build<T>(ctx, mb, wall);
    T* d_pts = ctx->d_pts.as<T>((size_t)N * 3);
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_pts, pts, (size_t)N * 3 * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    if (out_inside) {
        uint8_t* d_o = ctx->d_misc.as<uint8_t>((size_t)N);
        mesh_isinside<T>(ctx, mb, d_pts, N, d_o);
        WTP_CUDA_CHECK(cudaMemcpyAsync(out_inside, d_o, (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        T* d_o = ctx->d_misc.as<T>((size_t)N * 3);
        int64_t* d_t = ctx->d_misc2.as<int64_t>((size_t)N);
        mesh_project<T>(ctx, mb, d_pts, N, d_o, d_t);
        WTP_CUDA_CHECK(cudaMemcpyAsync(out_pts, d_o, (size_t)N * 3 * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
        WTP_CUDA_CHECK(cudaMemcpyAsync(out_tri, d_t, (size_t)N * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END(ctx)
}

// ------------------------------------------------- isinside(points, cloud)
template <class T>
static int32_t isinside_host(wtp_ctx* ctx, const T* pts, int64_t N, int32_t D, const T* bx, const T* bn, const T* ba, int64_t M,
                             uint8_t* out, T* sums) {
    API_BEGIN(ctx)
    WTP_REQUIRE(pts && bx && out && N > 0 && M > 0, WTP_ERR_BAD_ARG, "null pointer or empty point set");
    WTP_REQUIRE(D == 2 || D == 3, WTP_ERR_BAD_ARG, "D must be 2 or 3");
    WTP_REQUIRE(D == 2 || (bn && ba), WTP_ERR_BAD_ARG, "the 3-D test needs the boundary normals and areas");
    if (D == 2) {                                                    // _validate_polygon_ordering (src/isinside.jl:37-69)
        WTP_REQUIRE(M >= 3, WTP_ERR_BAD_ARG, "need at least 3 points to define a polygon");
        T sa = (T)0, xmin = bx[0], xmax = bx[0], ymin = bx[1], ymax = bx[1];
        for (int64_t i = 0; i < M; ++i) {
            const int64_t j = i + 1 == M ? 0 : i + 1;
            sa += bx[i * 2] * bx[j * 2 + 1] - bx[j * 2] * bx[i * 2 + 1];
            xmin = std::min(xmin, bx[i * 2]); xmax = std::max(xmax, bx[i * 2]);
            ymin = std::min(ymin, bx[i * 2 + 1]); ymax = std::max(ymax, bx[i * 2 + 1]);
        }
        sa /= (T)2;
        const T bbox_area = (xmax - xmin) * (ymax - ymin);
        WTP_REQUIRE(!(bbox_area > (T)0 && std::fabs(sa) < (T)1.0e-10 * bbox_area), WTP_ERR_BAD_ARG,
                    "polygon points do not appear to be ordered sequentially around the boundary");
    }
    T* d_pts = ctx->d_pts.as<T>((size_t)N * D);
    T* d_b = ctx->d_spacing_pts.as<T>((size_t)M * (D == 3 ? 7 : 2));
    uint8_t* d_o = ctx->d_misc.as<uint8_t>((size_t)N);
    T* d_s = sums ? ctx->d_misc2.as<T>((size_t)N) : nullptr;
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_pts, pts, (size_t)N * D * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_b, bx, (size_t)M * D * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    if (D == 3) {
        WTP_CUDA_CHECK(cudaMemcpyAsync(d_b + M * 3, bn, (size_t)M * 3 * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        WTP_CUDA_CHECK(cudaMemcpyAsync(d_b + M * 6, ba, (size_t)M * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        greens_isinside<T>(ctx, d_pts, N, d_b, d_b + M * 3, d_b + M * 6, M, d_s, d_o);
    } else {
        winding_isinside<T>(ctx, d_pts, N, d_b, M, d_s, d_o);
    }
    WTP_CUDA_CHECK(cudaMemcpyAsync(out, d_o, (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    if (sums) WTP_CUDA_CHECK(cudaMemcpyAsync(sums, d_s, (size_t)N * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END(ctx)
}

// ---------------------------------------------------------------- metrics
struct MetricsPartial { double avg, sd, mx, mn, sep, fill; };

// thread per point over its row of k sorted distances (rank 0 = self, dropped): the
// per-point mean/std/max/min of src/metrics.jl:22-30, block-reduced in a fixed order
template <class T>
__global__ void __launch_bounds__(256) metrics_kernel(const T* __restrict__ dist, int64_t N, int k, MetricsPartial* __restrict__ partials) {
    __shared__ MetricsPartial s[256];
    MetricsPartial acc{0, 0, 0, 0, 1.0e300, 0};
    const int m = k - 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const T* row = dist + i * k;
        T sum = (T)0;
        for (int j = 1; j < k; ++j) sum = sum + row[j];
        const T mean = sum / (T)m;
        T v = (T)0;
        for (int j = 1; j < k; ++j) { const T e = row[j] - mean; v = v + e * e; }
        const double sd = m > 1 ? (double)sqrt(v / (T)(m - 1)) : __longlong_as_double(0x7ff8000000000000LL);
        const double nn = (double)row[1];
        acc.avg += (double)mean; acc.sd += sd; acc.mx += (double)row[k - 1]; acc.mn += nn;
        acc.sep = nn < acc.sep ? nn : acc.sep;
        acc.fill = nn > acc.fill ? nn : acc.fill;
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            MetricsPartial a = s[threadIdx.x], b = s[threadIdx.x + o];
            a.avg += b.avg; a.sd += b.sd; a.mx += b.mx; a.mn += b.mn;
            a.sep = b.sep < a.sep ? b.sep : a.sep; a.fill = b.fill > a.fill ? b.fill : a.fill;
            s[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = s[0];
}

}  // namespace wtp

extern "C" {

int32_t wtp_repel_f32(wtp_ctx* c, float* snap, int64_t nf, int64_t nm, int32_t D, const wtp_spacing* sp, const wtp_force* fm,
                      const wtp_repel_params* prm, const wtp_wall_mesh* wall, float* conv, wtp_trace_entry* tr, wtp_repel_result* res) {
    return repel_entry<float>(c, snap, nf, nm, D, sp, fm, prm, wall, conv, tr, res);
}
int32_t wtp_repel_f64(wtp_ctx* c, double* snap, int64_t nf, int64_t nm, int32_t D, const wtp_spacing* sp, const wtp_force* fm,
                      const wtp_repel_params* prm, const wtp_wall_mesh* wall, double* conv, wtp_trace_entry* tr, wtp_repel_result* res) {
    return repel_entry<double>(c, snap, nf, nm, D, sp, fm, prm, wall, conv, tr, res);
}
int32_t wtp_repel_dev_f32(wtp_ctx* c, float* snap, int64_t nf, int64_t nm, int32_t D, const wtp_spacing* sp, const wtp_force* fm,
                          const wtp_repel_params* prm, float* conv, wtp_trace_entry* tr, wtp_repel_result* res) {
    return is_multi(c) ? no_device_pointers2(c) : repel_dev<float>(c, snap, nf, nm, D, sp, fm, prm, conv, tr, res);
}
int32_t wtp_repel_dev_f64(wtp_ctx* c, double* snap, int64_t nf, int64_t nm, int32_t D, const wtp_spacing* sp, const wtp_force* fm,
                          const wtp_repel_params* prm, double* conv, wtp_trace_entry* tr, wtp_repel_result* res) {
    return is_multi(c) ? no_device_pointers2(c) : repel_dev<double>(c, snap, nf, nm, D, sp, fm, prm, conv, tr, res);
}

int32_t wtp_spacing_eval_f32(wtp_ctx* c, const wtp_spacing* sp, const float* p, int64_t N, int32_t D, float* out) { return spacing_eval_host<float>(solo_of(c), sp, p, N, D, out); }
int32_t wtp_spacing_eval_f64(wtp_ctx* c, const wtp_spacing* sp, const double* p, int64_t N, int32_t D, double* out) { return spacing_eval_host<double>(solo_of(c), sp, p, N, D, out); }
int32_t wtp_force_eval_f32(wtp_ctx* c, const wtp_force* f, const float* u, int64_t n, float* out) { return force_eval_host<float>(solo_of(c), f, u, n, out); }
int32_t wtp_force_eval_f64(wtp_ctx* c, const wtp_force* f, const double* u, int64_t n, double* out) { return force_eval_host<double>(solo_of(c), f, u, n, out); }

int32_t wtp_isinside_f32(wtp_ctx* c, const float* p, int64_t N, int32_t D, const float* bx, const float* bn, const float* ba, int64_t M, uint8_t* out, float* sums) {
    return isinside_host<float>(solo_of(c), p, N, D, bx, bn, ba, M, out, sums);
}
int32_t wtp_isinside_f64(wtp_ctx* c, const double* p, int64_t N, int32_t D, const double* bx, const double* bn, const double* ba, int64_t M, uint8_t* out, double* sums) {
    return isinside_host<double>(solo_of(c), p, N, D, bx, bn, ba, M, out, sums);
}
int32_t wtp_mesh_isinside_f32(wtp_ctx* c, const wtp_wall_mesh* m, const float* p, int64_t N, uint8_t* out) { return mesh_query_host<float>(solo_of(c), m, p, N, out, nullptr, nullptr); }
int32_t wtp_mesh_isinside_f64(wtp_ctx* c, const wtp_wall_mesh* m, const double* p, int64_t N, uint8_t* out) { return mesh_query_host<double>(solo_of(c), m, p, N, out, nullptr, nullptr); }
int32_t wtp_mesh_project_f32(wtp_ctx* c, const wtp_wall_mesh* m, const float* p, int64_t N, float* op, int64_t* ot) { return mesh_query_host<float>(solo_of(c), m, p, N, nullptr, op, ot); }
int32_t wtp_mesh_project_f64(wtp_ctx* c, const wtp_wall_mesh* m, const double* p, int64_t N, double* op, int64_t* ot) { return mesh_query_host<double>(solo_of(c), m, p, N, nullptr, op, ot); }

}  // extern "C"

namespace wtp {
// defined in api.cu
template <class T>
void knn_device_self(wtp_ctx* ctx, const T* d_pts, int64_t N, int D, int k, int64_t* d_out_idx, T* d_out_dist);

template <class T>
static int32_t metrics_host(wtp_ctx* ctx, const T* pts, int64_t N, int32_t D, int32_t k, wtp_cloud_metrics* out) {
    API_BEGIN(ctx)
    WTP_REQUIRE(pts && out && N > 0, WTP_ERR_BAD_ARG, "null pointer or empty point set");
    WTP_REQUIRE(D == 2 || D == 3, WTP_ERR_BAD_ARG, "D must be 2 or 3");
    WTP_REQUIRE(k >= 2 && (int64_t)k <= N, WTP_ERR_K_TOO_LARGE, "metrics needs 2 <= k <= N");
    WTP_REQUIRE(ctx->world == 1, WTP_ERR_UNSUPPORTED, "metrics runs on a single-GPU context");
    T* d_pts = ctx->d_pts.as<T>((size_t)N * D);
    int64_t* d_idx = ctx->d_out_idx.as<int64_t>((size_t)N * k);
    T* d_dist = ctx->d_out_dist.as<T>((size_t)N * k);
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_pts, pts, (size_t)N * D * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    knn_device_self<T>(ctx, d_pts, N, D, k, d_idx, d_dist);
    const int nb = (int)std::min<int64_t>((N + 255) / 256, (int64_t)kNumSMs * 8);
    MetricsPartial* d_part = ctx->d_reduce.as<MetricsPartial>((size_t)nb);
    metrics_kernel<T><<<nb, 256, 0, ctx->stream>>>(d_dist, N, k, d_part);
    LAUNCH_CHECK(ctx);
    std::vector<MetricsPartial> h((size_t)nb);
    WTP_CUDA_CHECK(cudaMemcpyAsync(h.data(), d_part, sizeof(MetricsPartial) * nb, cudaMemcpyDeviceToHost, ctx->stream));
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    MetricsPartial a = h[0];
    for (int i = 1; i < nb; ++i) {
        a.avg += h[i].avg; a.sd += h[i].sd; a.mx += h[i].mx; a.mn += h[i].mn;
        a.sep = std::min(a.sep, h[i].sep); a.fill = std::max(a.fill, h[i].fill);
    }
    out->avg = a.avg / (double)N; out->std = a.sd / (double)N; out->max = a.mx / (double)N; out->min = a.mn / (double)N;
    out->separation = a.sep; out->fill = a.fill; out->mesh_ratio = a.fill / a.sep;
    API_END(ctx)
}
}  // namespace wtp

// ------------------------------------------------------------------ cull
namespace wtp {
template <class T>
static int32_t cull_mask_host(wtp_ctx* ctx, const T* pts, int64_t N, int32_t D, const T* spacings, double ratio, uint8_t* keep) {
    API_BEGIN(ctx)
    WTP_REQUIRE(pts && spacings && keep && N > 0, WTP_ERR_BAD_ARG, "null pointer or empty point set");
    WTP_REQUIRE(D == 2 || D == 3, WTP_ERR_BAD_ARG, "D must be 2 or 3");
    WTP_REQUIRE(ctx->world == 1, WTP_ERR_UNSUPPORTED, "the cull runs on a single-GPU context");
    memset(keep, 1, (size_t)N);
    if (!(ratio > 0) || N < 2) return WTP_OK;                                        // src/repel.jl:568
    T smax = spacings[0];
    for (int64_t i = 1; i < N; ++i) smax = std::max(smax, spacings[i]);
    const T r = (T)ratio * smax;                                                     // MetricBall(ratio * maximum(spacings)), :570
    // ball search on the device: CSR rows sorted by index, self removed
    ctx->timer.reset(ctx->stream);
    T* d_pts = ctx->d_pts.as<T>((size_t)N * D);
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_pts, pts, (size_t)N * D * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    int64_t* d_off = ctx->d_offsets.as<int64_t>((size_t)N + 1);
    radius_count_device<T>(ctx, d_pts, N, D, r, d_off);
    const int64_t nnz = ctx->radius.nnz;
    ctx->radius.pending = false;
    std::vector<int64_t> off((size_t)N + 1), ind((size_t)std::max<int64_t>(nnz, 1));
    WTP_CUDA_CHECK(cudaMemcpyAsync(off.data(), d_off, (size_t)(N + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (nnz > 0) {
        int64_t* d_ind = ctx->d_indices.as<int64_t>((size_t)nnz);
        radius_fill_device<T>(ctx, d_off, d_ind);
        WTP_CUDA_CHECK(cudaMemcpyAsync(ind.data(), d_ind, (size_t)nnz * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    // the greedy sweep in index order (:571-578): earlier decisions must be visible to later points
    for (int64_t i = 0; i < N; ++i) {
        if (!keep[i]) continue;
        const T thr = (T)ratio * spacings[i];
        for (int64_t e = off[(size_t)i]; e < off[(size_t)i + 1]; ++e) {
            const int64_t j = ind[(size_t)e] - 1;
            if (!keep[j]) continue;
            T d2 = (T)0;
            for (int d = 0; d < D; ++d) { const T v = pts[j * D + d] - pts[i * D + d]; d2 = d2 + v * v; }
            if (std::sqrt(d2) < thr) keep[j] = 0;
        }
    }
    API_END(ctx)
}
}  // namespace wtp
extern "C" {
int32_t wtp_cull_mask_f32(wtp_ctx* c, const float* p, int64_t N, int32_t D, const float* s, double ratio, uint8_t* keep) { return cull_mask_host<float>(solo_of(c), p, N, D, s, ratio, keep); }
int32_t wtp_cull_mask_f64(wtp_ctx* c, const double* p, int64_t N, int32_t D, const double* s, double ratio, uint8_t* keep) { return cull_mask_host<double>(solo_of(c), p, N, D, s, ratio, keep); }
}

// ------------------------------------------- spacing_metrics / spacing_fidelity_metrics
namespace wtp {

struct StatPartial { double sum, mx; };

// sum and max of (x[i] - shift)^power, power in {1, 2}: grid-stride, fixed block tree (mean / centred second moment)
template <class T>
__global__ void __launch_bounds__(256) stat_kernel(const T* __restrict__ x, int64_t n, double shift, int power, StatPartial* __restrict__ partials) {
    __shared__ StatPartial s[256];
    StatPartial acc{0.0, -1.0e300};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)x[i] - shift;
        acc.sum += power == 2 ? v * v : v;
        acc.mx = v > acc.mx ? v : acc.mx;
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { s[threadIdx.x].sum += s[threadIdx.x + o].sum; s[threadIdx.x].mx = fmax(s[threadIdx.x].mx, s[threadIdx.x + o].mx); }
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = s[0];
}

template <class T>
static StatPartial stat_reduce(wtp_ctx* ctx, const T* d_x, int64_t n, double shift, int power) {
    const int nb = (int)std::min<int64_t>((n + 255) / 256, (int64_t)kNumSMs * 8);
    StatPartial* d_part = ctx->d_reduce.as<StatPartial>((size_t)nb);
    stat_kernel<T><<<nb, 256, 0, ctx->stream>>>(d_x, n, shift, power, d_part);
    LAUNCH_CHECK(ctx);
    std::vector<StatPartial> h((size_t)nb);
    WTP_CUDA_CHECK(cudaMemcpyAsync(h.data(), d_part, sizeof(StatPartial) * nb, cudaMemcpyDeviceToHost, ctx->stream));
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    StatPartial a = h[0];
    for (int i = 1; i < nb; ++i) { a.sum += h[i].sum; a.mx = std::max(a.mx, h[i].mx); }
    return a;
}

// error_i = |mean(dist[i][1:]) - s_i| / s_i      (src/metrics.jl:61-63)
template <class T>
__global__ void __launch_bounds__(256) spacing_error_kernel(const T* __restrict__ dist, int64_t N, int k, const T* __restrict__ spacing, T* __restrict__ err) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const T* row = dist + i * k;
    T sum = (T)0;
    for (int j = 1; j < k; ++j) sum = sum + row[j];
    const T actual = sum / (T)(k - 1), target = spacing[i];
    err[i] = fabs(actual - target) / target;
}

// u_i = d_NN(i) / h_i with self skipped BY INDEX, coordination count within coord_radius * h_i   (src/metrics.jl:98-113)
template <class T>
__global__ void __launch_bounds__(256) fidelity_kernel(const int64_t* __restrict__ idx, const T* __restrict__ dist, int64_t N, int k,
                                                       const T* __restrict__ spacing, T coord_radius, T* __restrict__ u, float* __restrict__ coord) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const T h = spacing[i];
    T dmin = sizeof(T) == 4 ? (T)3.402823466e+38f : (T)1.7976931348623157e+308;
    int c = 0;
    for (int j = 0; j < k; ++j) {
        if (idx[i * k + j] == i + 1) continue;
        const T d = dist[i * k + j];
        dmin = d < dmin ? d : dmin;
        c += d <= coord_radius * h ? 1 : 0;
    }
    u[i] = dmin / h;
    coord[i] = (float)c;
}

// order statistics of non-negative values: radix sort of the bit patterns (two stable 32-bit sorts for doubles)
template <class T>
__global__ void __launch_bounds__(256) bits_key_kernel(const T* __restrict__ x, const uint32_t* __restrict__ order, int64_t n, int word,
                                                       uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t i = order ? order[t] : (uint32_t)t;
    uint32_t key;
    if (sizeof(T) == 4) key = __float_as_uint((float)x[i]);
    else { const unsigned long long b = (unsigned long long)__double_as_longlong((double)x[i]); key = word == 0 ? (uint32_t)b : (uint32_t)(b >> 32); }
    keys[t] = key;
    vals[t] = i;
}
template <class T>
__global__ void pick_kernel(const T* __restrict__ x, const uint32_t* __restrict__ order, const int64_t* __restrict__ pos, int npos, double* __restrict__ out) {
    if ((int)threadIdx.x < npos) out[threadIdx.x] = (double)x[order[pos[threadIdx.x]]];
}

// q[j] = quantile(x, p[j]) with linear interpolation between order statistics (Julia's default, type 7)
template <class T>
static void quantiles(wtp_ctx* ctx, const T* d_x, int64_t n, const double* p, int np, double* q) {
    IndexBuffers& ib = ctx->index[0];
    uint32_t* keys = ib.keys_a.as<uint32_t>((size_t)n);
    uint32_t* vals = ib.vals_a.as<uint32_t>((size_t)n);
    const unsigned nb = (unsigned)((n + 255) / 256);
    const int words = sizeof(T) == 4 ? 1 : 2;
    for (int w = 0; w < words; ++w) {
        // second word: keys are gathered in the order the first sort produced (LSD on 64 bits)
        const uint32_t* order = w == 0 ? nullptr : ctx->d_misc2.get<uint32_t>();
        bits_key_kernel<T><<<nb, 256, 0, ctx->stream>>>(d_x, order, n, w, keys, vals);
        LAUNCH_CHECK(ctx);
        radix_sort_pairs(ctx, ib, n, 32);
        vals = ib.vals_a.get<uint32_t>(); keys = ib.keys_a.get<uint32_t>();
        if (w + 1 < words) WTP_CUDA_CHECK(cudaMemcpyAsync(ctx->d_misc2.as<uint32_t>((size_t)n), vals, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    std::vector<int64_t> pos((size_t)2 * np);
    std::vector<double> frac((size_t)np);
    for (int j = 0; j < np; ++j) {
        const double hh = (double)(n - 1) * p[j];
        const int64_t lo = (int64_t)std::floor(hh);
        pos[2 * j] = lo; pos[2 * j + 1] = std::min<int64_t>(lo + 1, n - 1);
        frac[j] = hh - (double)lo;
    }
    int64_t* d_pos = ctx->d_offsets.as<int64_t>((size_t)2 * np + 2 * np);
    double* d_out = reinterpret_cast<double*>(d_pos + 2 * np);
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_pos, pos.data(), sizeof(int64_t) * 2 * np, cudaMemcpyHostToDevice, ctx->stream));
    pick_kernel<T><<<1, 32, 0, ctx->stream>>>(d_x, vals, d_pos, 2 * np, d_out);
    LAUNCH_CHECK(ctx);
    std::vector<double> v((size_t)2 * np);
    WTP_CUDA_CHECK(cudaMemcpyAsync(v.data(), d_out, sizeof(double) * 2 * np, cudaMemcpyDeviceToHost, ctx->stream));
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int j = 0; j < np; ++j) q[j] = v[2 * j] + frac[j] * (v[2 * j + 1] - v[2 * j]);
}

// uploads the cloud (and the spacing's boundary set), runs the k-NN with distances (k nearest including self) and
// evaluates the spacing at every point: the common front of the two spacing metrics
template <class T>
static void spacing_metrics_front(wtp_ctx* ctx, const T* pts, int64_t N, int D, int k, const wtp_spacing* sp, int64_t** d_idx, T** d_dist, T** d_spacing) {
    WTP_REQUIRE(sp->kind >= WTP_SPACING_CONSTANT && sp->kind <= WTP_SPACING_BOUNDARY_LAYER, WTP_ERR_UNSUPPORTED, "user-defined spacing callable cannot cross the C ABI");
    WTP_REQUIRE(ctx->world == 1, WTP_ERR_UNSUPPORTED, "metrics run on a single-GPU context");
    T* d_pts = ctx->d_pts.as<T>((size_t)N * D);
    *d_idx = ctx->d_out_idx.as<int64_t>((size_t)N * k);
    *d_dist = ctx->d_out_dist.as<T>((size_t)N * k);
    *d_spacing = ctx->d_spacings.as<T>((size_t)N);
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_pts, pts, (size_t)N * D * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    if (sp->kind != WTP_SPACING_CONSTANT) {
        WTP_REQUIRE(sp->bnd_pts && sp->n_bnd > 0, WTP_ERR_BAD_ARG, "variable spacing needs its boundary point set");
        T* b = ctx->d_spacing_pts.as<T>((size_t)sp->n_bnd * D);
        WTP_CUDA_CHECK(cudaMemcpyAsync(b, sp->bnd_pts, (size_t)sp->n_bnd * D * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        bvh_build<T>(ctx, ctx->bvh, b, sp->n_bnd, D);
    }
    const SpacingP<T> spp{sp->kind, (T)sp->a, (T)sp->b, (T)sp->c};
    spacing_eval<T>(ctx, spp, ctx->bvh, d_pts, N, D, *d_spacing);
    knn_device_self<T>(ctx, d_pts, N, D, k, *d_idx, *d_dist);
}

template <class T>
static int32_t spacing_metrics_host(wtp_ctx* ctx, const T* pts, int64_t N, int32_t D, int32_t k, const wtp_spacing* sp, wtp_spacing_metrics_t* out) {
    API_BEGIN(ctx)
    WTP_REQUIRE(pts && sp && out && N > 0, WTP_ERR_BAD_ARG, "null pointer or empty point set");
    WTP_REQUIRE(D == 2 || D == 3, WTP_ERR_BAD_ARG, "D must be 2 or 3");
    WTP_REQUIRE(k >= 2 && (int64_t)k <= N, WTP_ERR_K_TOO_LARGE, "spacing_metrics needs 2 <= k <= N");
    int64_t* d_idx; T* d_dist; T* d_s;
    spacing_metrics_front<T>(ctx, pts, N, D, k, sp, &d_idx, &d_dist, &d_s);
    T* d_err = ctx->d_nn.as<T>((size_t)N);
    spacing_error_kernel<T><<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_dist, N, k, d_s, d_err);
    LAUNCH_CHECK(ctx);
    const StatPartial a = stat_reduce<T>(ctx, d_err, N, 0.0, 1);
    const double mu = a.sum / (double)N;
    const StatPartial b = stat_reduce<T>(ctx, d_err, N, mu, 2);
    out->max_error = a.mx; out->mean_error = mu;
    out->std_error = N > 1 ? std::sqrt(b.sum / (double)(N - 1)) : std::nan("");
    API_END(ctx)
}

template <class T>
static int32_t spacing_fidelity_host(wtp_ctx* ctx, const T* pts, int64_t N, int32_t D, int32_t k, double coord_radius, const wtp_spacing* sp,
                                     wtp_spacing_fidelity_t* out) {
    API_BEGIN(ctx)
    WTP_REQUIRE(pts && sp && out && N > 0, WTP_ERR_BAD_ARG, "null pointer or empty point set");
    WTP_REQUIRE(D == 2 || D == 3, WTP_ERR_BAD_ARG, "D must be 2 or 3");
    k = (int32_t)std::min<int64_t>(k, N);                                                    // src/metrics.jl:93
    WTP_REQUIRE(k >= 1 && k <= WTP_MAX_K, WTP_ERR_K_TOO_LARGE, "k exceeds WTP_MAX_K");
    int64_t* d_idx; T* d_dist; T* d_s;
    spacing_metrics_front<T>(ctx, pts, N, D, k, sp, &d_idx, &d_dist, &d_s);
    T* d_u = ctx->d_nn.as<T>((size_t)N);
    float* d_c = ctx->d_counts.as<float>((size_t)N);
    fidelity_kernel<T><<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_idx, d_dist, N, k, d_s, (T)coord_radius, d_u, d_c);
    LAUNCH_CHECK(ctx);
    const StatPartial a = stat_reduce<T>(ctx, d_u, N, 0.0, 1);
    const double mu = a.sum / (double)N;
    const StatPartial b = stat_reduce<T>(ctx, d_u, N, mu, 2);
    const StatPartial c = stat_reduce<float>(ctx, d_c, N, 0.0, 1);
    const double p[3] = {0.05, 0.5, 0.95};
    double q[3];
    quantiles<T>(ctx, d_u, N, p, 3, q);
    out->mean_dnn_h = mu;
    out->cv = (N > 1 ? std::sqrt(b.sum / (double)(N - 1)) : std::nan("")) / mu;
    out->p05 = q[0]; out->p50 = q[1]; out->p95 = q[2];
    out->coordination = c.sum / (double)N;
    API_END(ctx)
}

}  // namespace wtp

extern "C" {
int32_t wtp_spacing_metrics_f32(wtp_ctx* c, const float* p, int64_t N, int32_t D, int32_t k, const wtp_spacing* sp, wtp_spacing_metrics_t* out) { return spacing_metrics_host<float>(solo_of(c), p, N, D, k, sp, out); }
int32_t wtp_spacing_metrics_f64(wtp_ctx* c, const double* p, int64_t N, int32_t D, int32_t k, const wtp_spacing* sp, wtp_spacing_metrics_t* out) { return spacing_metrics_host<double>(solo_of(c), p, N, D, k, sp, out); }
int32_t wtp_spacing_fidelity_f32(wtp_ctx* c, const float* p, int64_t N, int32_t D, int32_t k, double cr, const wtp_spacing* sp, wtp_spacing_fidelity_t* out) { return spacing_fidelity_host<float>(solo_of(c), p, N, D, k, cr, sp, out); }
int32_t wtp_spacing_fidelity_f64(wtp_ctx* c, const double* p, int64_t N, int32_t D, int32_t k, double cr, const wtp_spacing* sp, wtp_spacing_fidelity_t* out) { return spacing_fidelity_host<double>(solo_of(c), p, N, D, k, cr, sp, out); }
}

extern "C" {
int32_t wtp_metrics_f32(wtp_ctx* c, const float* p, int64_t N, int32_t D, int32_t k, wtp_cloud_metrics* out) { return metrics_host<float>(solo_of(c), p, N, D, k, out); }
int32_t wtp_metrics_f64(wtp_ctx* c, const double* p, int64_t N, int32_t D, int32_t k, wtp_cloud_metrics* out) { return metrics_host<double>(solo_of(c), p, N, D, k, out); }
}
