// api.cu — the C ABI of libwtp_cuda.so (include/wtp_cuda.h): context, instrumentation,
// and the topology entry points. Repel lives in repel.cu, multi-GPU plumbing in comm.cu.
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <new>

#include "d2h_pipeline.h"
#include "host_pool.h"
#include "kernels.cuh"
#include "multi.h"

using namespace wtp;

namespace wtp {
static std::atomic<uint64_t> g_error_stamp{0};
uint64_t next_error_stamp() { return ++g_error_stamp; }
int32_t fail(wtp_ctx* ctx, const Error& e) {
    if (ctx) { ctx->last_error = e.msg; ctx->error_stamp = next_error_stamp(); }
    return e.status;
}
void finish_timing(wtp_ctx* ctx, int sort_passes, int query_launches, int64_t n_cells, int64_t n_expanded);
}  // namespace wtp

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

extern "C" {

int32_t wtp_version(void) { return 100; }

const char* wtp_status_string(int32_t s) {
    switch (s) {
        case WTP_OK: return "ok";
        case WTP_ERR_BAD_ARG: return "bad argument";
        case WTP_ERR_K_TOO_LARGE: return "k too large for the point set or above WTP_MAX_K";
        case WTP_ERR_UNSUPPORTED: return "unsupported force model / spacing / option (no CPU fallback)";
        case WTP_ERR_CUDA: return "CUDA error";
        case WTP_ERR_NCCL: return "NCCL error";
        case WTP_ERR_OOM: return "out of memory";
        case WTP_ERR_STATE: return "call sequence error";
    }
    return "unknown status";
}

void wtp_comm_destroy_internal(wtp_ctx* ctx);

int32_t wtp_create(wtp_ctx** out, int32_t device) {
    if (!out) return WTP_ERR_BAD_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) { (void)cudaGetLastError(); return WTP_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return WTP_ERR_CUDA;
    if (prop.major < 10) return WTP_ERR_CUDA;  // sm_100a code only
    wtp_ctx* ctx = new (std::nothrow) wtp_ctx();
    if (!ctx) return WTP_ERR_OOM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return WTP_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    ctx->h_pinned_bytes = 4096;
    if (cudaMallocHost(&ctx->h_pinned, ctx->h_pinned_bytes) != cudaSuccess) { cudaStreamDestroy(ctx->own_stream); delete ctx; return WTP_ERR_CUDA; }
    bool ok = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (auto& e : ctx->ev_chunk) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_copy_done, cudaEventDisableTiming) == cudaSuccess;
    for (auto& e : ctx->ev_copied) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { wtp_destroy(ctx); return WTP_ERR_CUDA; }
    *out = ctx;
    return WTP_OK;
}

void wtp_destroy(wtp_ctx* ctx) {
    if (!ctx) return;
    if (!ctx->children.empty() || ctx->solo) {   // the parent of a multi-device context: no device state of its own
        for (wtp_ctx* c : ctx->children) { c->group = nullptr; wtp_destroy(c); }
        if (ctx->solo) wtp_destroy(ctx->solo);
        delete ctx->group;
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    wtp_comm_destroy_internal(ctx);
    delete ctx->pool;
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    for (auto e : ctx->ev_copied) if (e) cudaEventDestroy(e);
    for (auto e : ctx->ev_ring) if (e) cudaEventDestroy(e);
    if (ctx->h_ids) cudaFreeHost(ctx->h_ids);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (auto e : ctx->ev_chunk) if (e) cudaEventDestroy(e);
    if (ctx->ev_copy_done) cudaEventDestroy(ctx->ev_copy_done);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* wtp_last_error(const wtp_ctx* ctx) {
    if (!ctx) return "null context";
    const wtp_ctx* latest = ctx;                   // a multi-device context: the most recent failure of the parent or its single-device context
    if (ctx->solo && ctx->solo->error_stamp > latest->error_stamp) latest = ctx->solo;
    return latest->last_error.c_str();
}

int32_t wtp_set_stream(wtp_ctx* ctx, void* s) {
    if (!ctx) return WTP_ERR_BAD_ARG;
    if (is_multi(ctx)) return fail(ctx, Error{WTP_ERR_UNSUPPORTED, "a multi-device context runs on its own streams (one per device)"});
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return WTP_OK;
}

int32_t wtp_host_register(wtp_ctx* ctx, void* ptr, int64_t bytes) {
    ctx = solo_of(ctx);
    API_BEGIN(ctx)
    WTP_REQUIRE(ptr && bytes > 0, WTP_ERR_BAD_ARG, "wtp_host_register: null range");
    WTP_CUDA_CHECK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));   // every device of the process may copy to / from it
    API_END(ctx)
}
int32_t wtp_host_unregister(wtp_ctx* ctx, void* ptr) {
    ctx = solo_of(ctx);
    API_BEGIN(ctx)
    WTP_CUDA_CHECK(cudaHostUnregister(ptr));
    API_END(ctx)
}

int32_t wtp_set_cell_occupancy(wtp_ctx* ctx, double m) {
    if (!ctx) return WTP_ERR_BAD_ARG;
    for (wtp_ctx* c : ctx->children) c->cell_occupancy = m;
    if (ctx->solo) ctx->solo->cell_occupancy = m;
    ctx->cell_occupancy = m;
    return WTP_OK;
}

int32_t wtp_set_timing(wtp_ctx* ctx, int32_t enable) {
    if (!ctx) return WTP_ERR_BAD_ARG;
    for (wtp_ctx* c : ctx->children) c->timer.enabled = enable != 0;
    if (ctx->solo) ctx->solo->timer.enabled = enable != 0;
    ctx->timer.enabled = enable != 0;
    return WTP_OK;
}

int32_t wtp_get_timing(wtp_ctx* ctx, wtp_timing* out) {
    if (is_multi(ctx)) ctx = ctx->children[0];     // rank 0's view of the last sharded call
    API_BEGIN(ctx)
    WTP_REQUIRE(out, WTP_ERR_BAD_ARG, "null output");
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    wtp_timing t = ctx->last_timing;
    float* slot[PH_COUNT] = {&t.ms_h2d, &t.ms_bbox, &t.ms_cellkey, &t.ms_sort, &t.ms_reorder, &t.ms_query,
                             &t.ms_scan, &t.ms_reduce, &t.ms_comm, &t.ms_d2h};
    for (int i = 0; i < PH_COUNT; ++i) *slot[i] = 0.f;
    t.ms_total = 0.f;
    for (auto& s : ctx->timer.spans) {
        if (!s.b) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) *slot[s.phase] += ms;
    }
    if (ctx->timer.have_total) cudaEventElapsedTime(&t.ms_total, ctx->timer.total_a, ctx->timer.total_b);
    *out = t;
    API_END(ctx)
}

int64_t wtp_launch_count(const wtp_ctx* ctx) {
    if (!ctx) return -1;
    int64_t n = ctx->launches;
    for (const wtp_ctx* c : ctx->children) n += c->launches;
    if (ctx->solo) n += ctx->solo->launches;
    return n;
}

int64_t wtp_shard_begin(int64_t n, int32_t rank, int32_t world) {
    if (world <= 1) return 0;
    return (n * (int64_t)rank) / world;
}
int64_t wtp_shard_end(int64_t n, int32_t rank, int32_t world) {
    if (world <= 1) return n;
    return (n * (int64_t)(rank + 1)) / world;
}

}  // extern "C"

namespace wtp {

void finish_timing(wtp_ctx* ctx, int sort_passes, int query_launches, int64_t n_cells, int64_t n_expanded) {
    ctx->last_timing = wtp_timing{};
    ctx->last_timing.sort_passes = sort_passes;
    ctx->last_timing.query_launches = query_launches;
    ctx->last_timing.n_cells = n_cells;
    ctx->last_timing.n_ring_expanded = n_expanded;
    ctx->last_timing.n_leftover_sparse = ctx->last_tile_sparse;
    ctx->last_timing.n_leftover_dense = ctx->last_tile_dense;
    ctx->last_timing.n_leftover_other = ctx->last_tile_other;
    ctx->last_timing.n_window_points = ctx->last_window_points;
    ctx->last_timing.n_window_missed = ctx->last_window_missed;
}

// The pinned staging ring of the host entry points (d2h_pipeline.h): R slots of WTP_STAGE_MB MiB (default 4, R = 8).
// Small slots keep the freshly DMA-written data in the last-level cache until the widening threads read it; the ring is
// deep enough to keep the link busy while they do. (Re)allocated when the geometry changes.
static size_t ensure_stage(wtp_ctx* ctx) {
    size_t mb = 4;
    if (const char* e = std::getenv("WTP_STAGE_MB")) { const long v = std::atol(e); if (v >= 1 && v <= 256) mb = (size_t)v; }
    int ring = (int)std::max<size_t>(3, std::min<size_t>(16, 32 / mb));
    if (const char* e = std::getenv("WTP_STAGE_SLOTS")) { const long v = std::atol(e); if (v >= 2 && v <= 64) ring = (int)v; }
    const size_t slot = mb << 20;
    if (!ctx->h_stage || ctx->h_stage_slot_bytes != slot || ctx->h_stage_ring != ring) {
        if (ctx->h_stage) { cudaFreeHost(ctx->h_stage); ctx->h_stage = nullptr; }
        WTP_CUDA_CHECK(cudaMallocHost(&ctx->h_stage, (size_t)ring * slot));
        ctx->h_stage_slot_bytes = slot;
        ctx->h_stage_ring = ring;
    }
    // one of the pool's threads is the pipeline's producer and sleeps most of the time (d2h_pipeline.h): when the host is
    // shared out among ranks it comes on top of the rank's share of the cores
    if (!ctx->pool) ctx->pool = new HostPool(HostPool::default_threads(ctx->world) + (ctx->world > 1 ? 1 : 0));
    return slot;
}

// 4 indices below 2^24 -> 12 bytes (3 little-endian bytes each)
__global__ void __launch_bounds__(256) pack24_kernel(const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ out) {
    const size_t groups = (n + 3) / 4;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
        uint32_t v[4];
        if (4 * g + 4 <= n) { const uint4 q = reinterpret_cast<const uint4*>(in)[g]; v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
        else for (int j = 0; j < 4; ++j) v[j] = 4 * g + j < n ? in[4 * g + j] : 0u;
        out[3 * g + 0] = v[0] | (v[1] << 24);
        out[3 * g + 1] = (v[1] >> 8) | (v[2] << 16);
        out[3 * g + 2] = (v[2] >> 16) | (v[3] << 8);
    }
}

// A device array of 4-byte indices (16-byte aligned) -> the caller's int64 host array: chunk by chunk through the
// staging ring on the copy stream, widened by the host pool while the next chunks are on the wire, so the caller's array
// does not have to be pinned and at most 4 bytes per entry cross PCIe instead of 8 — 3 when every value is below 2^24
// (max_value; point sets of up to 16.7 M): the link, not the host, is what bounds this step, so the entries are packed
// to 3 bytes on the device first (WTP_NO_PACK24 turns that off). The data must be complete on ctx->stream.
void d2h_widen_u32(wtp_ctx* ctx, const uint32_t* d_src, size_t n, int64_t* h_dst, uint64_t max_value) {
    if (n == 0) return;
    const size_t slot = ensure_stage(ctx);
    if (max_value < ((uint64_t)1 << 24) && std::getenv("WTP_NO_PACK24") == nullptr) {
        const size_t groups = (n + 3) / 4;
        uint32_t* d_packed = ctx->d_pack.as<uint32_t>(3 * groups + 4);
        const unsigned nb = (unsigned)std::min<size_t>((groups + 255) / 256, (size_t)kNumSMs * 16);
        pack24_kernel<<<nb, 256, 0, ctx->stream>>>(d_src, n, d_packed);
        ctx->launches++;
        WTP_CUDA_CHECK(cudaPeekAtLastError());
        const size_t chunk = ((slot - 16) / 12) * 12;                        // whole groups, 4 readable bytes behind the last one
        d2h_pipeline(ctx, d_packed, 12 * groups, chunk, slot, ctx->h_stage_ring,
                     [&](int64_t, const char* staged, size_t off, size_t len, int w, int workers) {
                         const size_t first = off / 12 * 4, total = std::min(len / 12 * 4, n - first);
                         const size_t a = (total * (size_t)w / (size_t)workers) & ~(size_t)3;
                         const size_t b = w + 1 == workers ? total : ((total * (size_t)(w + 1) / (size_t)workers) & ~(size_t)3);
                         widen_u24_to_i64(reinterpret_cast<const unsigned char*>(staged) + 3 * a, h_dst + first + a, b - a);
                     });
        ctx->last_d2h_bytes = 12 * groups;
        return;
    }
    d2h_pipeline(ctx, d_src, n * sizeof(uint32_t), slot, slot, ctx->h_stage_ring,
                 [&](int64_t, const char* staged, size_t off, size_t len, int w, int workers) {
                     const size_t total = len / sizeof(uint32_t), first = off / sizeof(uint32_t);
                     const size_t a = (total * (size_t)w / (size_t)workers) & ~(size_t)3;
                     const size_t b = w + 1 == workers ? total : ((total * (size_t)(w + 1) / (size_t)workers) & ~(size_t)3);
                     widen_u32_to_i64(reinterpret_cast<const uint32_t*>(staged) + a, h_dst + first + a, b - a);
                 });
    ctx->last_d2h_bytes = n * sizeof(uint32_t);
}

bool comm_peer_buffers(wtp_ctx* ctx, PeerSet& pb, size_t bytes_each);                                   // comm.cu
void comm_allgather_rows(wtp_ctx* ctx, void* d_buf, int64_t n_rows, size_t row_bytes);
void comm_allgather_fixed(wtp_ctx* ctx, const void* d_in, void* d_out, size_t bytes_per_rank);

__global__ void __launch_bounds__(256) widen_u32_kernel(const uint32_t* __restrict__ in, size_t n, int64_t* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (int64_t)in[i];
}
static void widen_on_device(wtp_ctx* ctx, const uint32_t* d_in, size_t n, int64_t* d_out) {
    if (n == 0) return;
    const unsigned nb = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)kNumSMs * 16);
    widen_u32_kernel<<<nb, 256, 0, ctx->stream>>>(d_in, n, d_out);
    ctx->launches++;
    WTP_CUDA_CHECK(cudaPeekAtLastError());
}

// ------------------------------------------------------------------- k-NN
// Index build + queries. Single context: every point is a query and rows go by caller index. Sharded context
// (wtp_comm_init): this rank answers the contiguous run [sb, se) of the spatially sorted order (the sorted set is
// replicated, no collective) and writes a compact table, row t = sorted position sb + t; wtp_shard_owned tells
// which caller index each row belongs to.
//   device sink (h_out_idx == null): d_out_idx / d_out_dist are the caller's device tables (int64 / T);
//   host sink: the rows are computed as 4-byte indices into a context buffer, brought back chunk by chunk into a
//   pinned staging ring on the copy stream and widened (and, when sharded, scattered by caller index) into the
//   caller's int64 table on the host pool while the next chunk is on the wire: 4 B per neighbour cross PCIe, not 8.
template <class T>
static void knn_device(wtp_ctx* ctx, const T* d_pts, int64_t N, int D, int k, bool drop_first, int64_t* d_out_idx, T* d_out_dist,
                       int64_t* h_out_idx = nullptr, T* h_out_dist = nullptr, bool dev_out32 = false) {
    WTP_REQUIRE(D == 2 || D == 3, WTP_ERR_BAD_ARG, "D must be 2 or 3");
    WTP_REQUIRE(k >= 1, WTP_ERR_BAD_ARG, "k must be >= 1");
    const int K1 = k + (drop_first ? 1 : 0);
    WTP_REQUIRE((int64_t)K1 <= N, WTP_ERR_K_TOO_LARGE, "k-nearest search needs at least k+1 points (k <= N for search)");
    WTP_REQUIRE(K1 <= WTP_MAX_K, WTP_ERR_K_TOO_LARGE, "k exceeds WTP_MAX_K");
    IndexBuffers& ib = ctx->index[0];
    double lo[3], hi[3];
    compute_bbox<T>(ctx, ib, d_pts, N, D, lo, hi);
    Grid<T> g = make_grid<T>(N, D, lo, hi, ctx->cell_occupancy, 0.0, K1);
    const bool sharded = ctx->world > 1;
    // CTA-tiled front end (knn_tile.cuh) whenever the list fits one register row, the general kernel alone otherwise
    const bool tiled = K1 <= 32 && std::getenv("WTP_NO_TILED") == nullptr;
    // the run of the sorted order this rank answers, in positions of the whole sorted set
    const int64_t gsb = wtp_shard_begin(N, ctx->rank, ctx->world), gse = wtp_shard_end(N, ctx->rank, ctx->world);
    const int64_t nq = gse - gsb;
    // Sharded: index only the layers of the grid that the run needs (build_index_window); sb, se are then positions
    // inside the window. A query whose search leaves the window is counted (d_exp[1]) and the call is repeated on
    // the whole index; the context remembers that and builds the whole index at once from then on.
    IndexWindow win;
    int passes = 0;
    bool windowed = sharded && tiled && !ctx->window_off && std::getenv("WTP_NO_WINDOW") == nullptr &&
                    build_index_window<T>(ctx, ib, d_pts, N, D, g, gsb, gse, 2, &win, &passes);
    if (!windowed) passes = build_index<T>(ctx, ib, d_pts, N, D, g);
    int64_t sb = windowed ? gsb - win.P0 : gsb, se = windowed ? gse - win.P0 : gse;
    ctx->owned_begin = sb; ctx->owned_end = se; ctx->owned_f64 = sizeof(T) == 8;
    ctx->last_window_points = windowed ? win.M : 0; ctx->last_window_missed = 0;
    RowMap rows{0u, (uint32_t)sb, sharded ? 1 : 0};
    unsigned long long* d_exp = ctx->d_reduce.as<unsigned long long>(2);
    WTP_CUDA_CHECK(cudaMemsetAsync(d_exp, 0, 2 * sizeof(unsigned long long), ctx->stream));
    auto compute_once = [&](void* d_idx, T* d_dist, bool out32) {
        if (tiled) knn_query_tiled<T>(ctx, ib, g, N, D, K1, drop_first ? 1 : 0, sb, se, rows, d_idx, d_dist, d_exp, out32);
        else knn_query<T>(ctx, ib, g, N, D, K1, drop_first ? 1 : 0, nullptr, nq, rows, d_idx, d_dist, d_exp, out32);
    };
    // h_pinned: [0] ring-expanded queries, [1] window misses, then the four counters of the tiled pass
    unsigned long long* h_exp = static_cast<unsigned long long*>(ctx->h_pinned);
    uint32_t* h_cnt = reinterpret_cast<uint32_t*>(h_exp + 2);
    bool counters_read = false;
    auto read_counters = [&] {
        WTP_CUDA_CHECK(cudaMemcpyAsync(h_exp, d_exp, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        if (tiled) WTP_CUDA_CHECK(cudaMemcpyAsync(h_cnt, ctx->d_fail.get<uint32_t>(), 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    };
    auto compute = [&](void* d_idx, T* d_dist, bool out32) {
        compute_once(d_idx, d_dist, out32);
        if (!windowed) return;
        read_counters();
        counters_read = true;
        if (h_exp[1] == 0) return;
        ctx->last_window_missed = (int64_t)h_exp[1];
        ctx->window_off = true;
        windowed = false;
        counters_read = false;
        g.w_lo = 0; g.w_hi = g.n[D - 1] - 1;
        passes = build_index<T>(ctx, ib, d_pts, N, D, g);
        sb = gsb; se = gse;
        ctx->owned_begin = sb; ctx->owned_end = se;
        rows.pos_base = (uint32_t)sb;
        WTP_CUDA_CHECK(cudaMemsetAsync(d_exp, 0, 2 * sizeof(unsigned long long), ctx->stream));
        compute_once(d_idx, d_dist, out32);
    };
    int n_chunks = 1;
    ctx->owned_contiguous = false;
    // Sharded host call on a communicator whose ranks can map each other's memory: the row exchange. The k-NN kernels store
    // the row of caller index i straight into the buffer of the rank that owns the caller range of i (peer memory over
    // NVLink, RowMap::elem), so that after one barrier every rank holds the rows of ONE CONTIGUOUS part of the caller's
    // table: it goes back with one forward copy — as int64 written by the DMA engine itself when the caller's table is
    // pinned and WTP_D2H_DIRECT=1, through the widening pipeline by default (see below) — instead of rows scattered all
    // over the table by host threads.
    const int64_t cb = wtp_shard_begin(N, ctx->rank, ctx->world), ce = wtp_shard_end(N, ctx->rank, ctx->world);   // caller range of this rank
    const size_t rx_each = (size_t)((N + ctx->world - 1) / ctx->world) * (size_t)k * sizeof(uint32_t);
    const bool exchange = sharded && h_out_idx && !h_out_dist && ctx->nccl_comm && std::getenv("WTP_NO_ROW_EXCHANGE") == nullptr &&
                          comm_peer_buffers(ctx, ctx->row_peers, rx_each);
    if (exchange) {
        const size_t parity = (size_t)(ctx->row_exchanges++ & 1u);
        auto rx = [&](int r) { return reinterpret_cast<uint32_t*>(static_cast<char*>(ctx->row_peers.base[r]) + parity * ctx->row_peers.bytes_each); };
        uint32_t* mine = rx(ctx->rank);
        rows.n_owners = ctx->world;
        for (int r = 0; r <= ctx->world; ++r) rows.owner_begin[r] = (uint32_t)wtp_shard_begin(N, r, ctx->world);
        rows.owner_begin[ctx->world] = (uint32_t)N;
        for (int r = 0; r < ctx->world; ++r)
            rows.owner_delta[r] = (long long)(rx(r) - mine) - (long long)rows.owner_begin[r] * k;
        compute(mine, nullptr, true);
        {   // barrier: every rank's kernels have finished storing into this rank's buffer
            ScopedPhase ph(ctx->timer, PH_COMM);
            uint64_t* d_flag = ctx->d_misc.as<uint64_t>((size_t)ctx->world + 1);
            comm_allgather_fixed(ctx, d_flag + ctx->world, d_flag, sizeof(uint64_t));
        }
        ScopedPhase ph(ctx->timer, PH_D2H);
        const size_t n_elems = (size_t)(ce - cb) * k;
        int64_t* h_dst = h_out_idx + cb * k;
        cudaPointerAttributes attr{};
        const bool pinned = n_elems > 0 && cudaPointerGetAttributes(&attr, h_dst) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        (void)cudaGetLastError();
        // Opt-in (WTP_D2H_DIRECT=1): measured on the 8-GPU box, DMA writes into host memory are capped at ~92 GB/s for
        // all GPUs together — two links' worth — so 8 bytes per entry written by the copy engines (18 ms for the 1.68 GB
        // table, at 2 GPUs and at 8) lose to 3 or 4 bytes per entry through the staging ring with the host threads widening.
        const char* force = std::getenv("WTP_D2H_DIRECT");
        const bool direct = force && force[0] == '1' && pinned;
        if (direct) {
            int64_t* d_wide = ctx->d_indices.as<int64_t>(std::max<size_t>(n_elems, 1));
            widen_on_device(ctx, mine, n_elems, d_wide);
            WTP_CUDA_CHECK(cudaMemcpyAsync(h_dst, d_wide, n_elems * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            d2h_widen_u32(ctx, mine, n_elems, h_dst, (uint64_t)N);
        }
        ctx->owned_contiguous = true;
        ctx->owned_begin = cb; ctx->owned_end = ce;
        ctx->last_d2h_direct = direct;
    } else if (!h_out_idx) {
        compute(d_out_idx, d_out_dist, dev_out32);    // dev_out32: d_out_idx is an int32 table (wtp_knn_dev_i32_*)
    } else if (!sharded && nq * (int64_t)k < ((int64_t)4 << 20)) {
        // small result: int64 rows straight from the device table
        int64_t* d_idx = ctx->d_out_idx.as<int64_t>((size_t)nq * k);
        T* d_dist = h_out_dist ? ctx->d_out_dist.as<T>((size_t)nq * k) : nullptr;
        compute(d_idx, d_dist, false);
        ScopedPhase ph(ctx->timer, PH_D2H);
        WTP_CUDA_CHECK(cudaMemcpyAsync(h_out_idx, d_idx, (size_t)nq * k * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (h_out_dist) WTP_CUDA_CHECK(cudaMemcpyAsync(h_out_dist, d_dist, (size_t)nq * k * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        const size_t SLOT = ensure_stage(ctx);
        uint32_t* d_idx32 = ctx->d_out_idx.as<uint32_t>((size_t)nq * k);
        T* d_dist = h_out_dist ? ctx->d_out_dist.as<T>((size_t)nq * k) : nullptr;
        compute(d_idx32, d_dist, true);
        // sharded: row t belongs to caller index ids[t] (1-based; 4-byte values in a pinned buffer of the context)
        const uint32_t* ids = nullptr;
        std::vector<T> dist_tmp;
        if (sharded) {
            const size_t need = (size_t)std::max<int64_t>(nq, 1) * sizeof(uint32_t);
            if (need > ctx->h_ids_bytes) {
                if (ctx->h_ids) cudaFreeHost(ctx->h_ids);
                ctx->h_ids = nullptr; ctx->h_ids_bytes = 0;
                WTP_CUDA_CHECK(cudaMallocHost(&ctx->h_ids, need + need / 8));
                ctx->h_ids_bytes = need + need / 8;
            }
            // the rows in ascending caller index: the host then writes forward through the caller's table
            uint32_t* d_ids = ctx->d_misc.as<uint32_t>((size_t)std::max<int64_t>(nq, 1));
            uint32_t* d_rows2 = ctx->d_indices.as<uint32_t>((size_t)nq * k);
            T* d_dist2 = h_out_dist ? ctx->d_misc2.as<T>((size_t)nq * k) : nullptr;
            rows_by_caller_index<T>(ctx, ib, N, sb, nq, k, d_idx32, d_dist, d_rows2, d_dist2, d_ids);
            d_idx32 = d_rows2;
            if (h_out_dist) d_dist = d_dist2;
            WTP_CUDA_CHECK(cudaMemcpyAsync(ctx->h_ids, d_ids, (size_t)nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            ids = static_cast<const uint32_t*>(ctx->h_ids);
            if (h_out_dist) {
                dist_tmp.resize((size_t)nq * k);
                WTP_CUDA_CHECK(cudaMemcpyAsync(dist_tmp.data(), d_dist, (size_t)nq * k * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
            }
            WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));               // ids (and distances) are on the host
        } else if (h_out_dist) {
            WTP_CUDA_CHECK(cudaMemcpyAsync(h_out_dist, d_dist, (size_t)nq * k * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
        }
        const size_t row_bytes = (size_t)k * sizeof(uint32_t);
        const size_t chunk_bytes = std::max<size_t>(1, SLOT / row_bytes) * row_bytes;   // whole rows per chunk
        n_chunks = (int)(((size_t)nq * row_bytes + chunk_bytes - 1) / chunk_bytes);
        const auto t_pipe = std::chrono::steady_clock::now();
        if (!sharded) {
            d2h_widen_u32(ctx, d_idx32, (size_t)nq * k, h_out_idx, (uint64_t)N);
        } else {
            d2h_pipeline(ctx, d_idx32, (size_t)nq * row_bytes, chunk_bytes, SLOT, ctx->h_stage_ring,
                         [&](int64_t, const char* staged, size_t off, size_t len, int w, int workers) {
                             const uint32_t* src = reinterpret_cast<const uint32_t*>(staged);
                             const int64_t nrow = (int64_t)(len / row_bytes), cb = (int64_t)(off / row_bytes);
                             const int64_t t_end = nrow * (w + 1) / workers;
                             for (int64_t t = nrow * w / workers; t < t_end; ++t) {
                                 int64_t* dst = h_out_idx + ((int64_t)ids[(size_t)(cb + t)] - 1) * k;
                                 for (int r = 0; r < k; ++r) _mm_stream_si64(reinterpret_cast<long long*>(dst + r), (long long)src[t * k + r]);   // no read-for-ownership of the caller's table
                                 if (h_out_dist) memcpy(h_out_dist + ((int64_t)ids[(size_t)(cb + t)] - 1) * k, dist_tmp.data() + (size_t)(cb + t) * k, (size_t)k * sizeof(T));
                             }
                             _mm_sfence();
                         });
            ctx->last_d2h_bytes = (size_t)nq * row_bytes + (size_t)nq * 4;
        }
        if (std::getenv("WTP_PIPE_DEBUG"))
            fprintf(stderr, "[wtp pipe] chunks=%d slot=%zu MiB ring=%d threads=%d d2h+widen=%.2f ms\n", n_chunks, SLOT >> 20, ctx->h_stage_ring, ctx->pool->size(),
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_pipe).count());
    }
    if (!counters_read || h_out_idx) {
        if (!counters_read) read_counters();
        else WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->last_tile_sparse = tiled ? h_cnt[1] : 0; ctx->last_tile_dense = tiled ? h_cnt[2] + h_cnt[4] : 0; ctx->last_tile_other = tiled ? h_cnt[3] : 0;
    finish_timing(ctx, passes, n_chunks, g.ncells, (int64_t)*h_exp);
    ctx->last_timing.n_peer_ranks = exchange ? ctx->world : 0;
    ctx->last_timing.bytes_d2h = !h_out_idx ? 0
        : exchange && ctx->last_d2h_direct ? (ce - cb) * (int64_t)k * 8
        : (!sharded && nq * (int64_t)k < ((int64_t)4 << 20)) ? nq * (int64_t)k * 8
        : (int64_t)ctx->last_d2h_bytes;
    if (h_out_idx && h_out_dist) ctx->last_timing.bytes_d2h += nq * (int64_t)k * (int64_t)sizeof(T);
}

template <class T>
void knn_device_self(wtp_ctx* ctx, const T* d_pts, int64_t N, int D, int k, int64_t* d_out_idx, T* d_out_dist) {
    knn_device<T>(ctx, d_pts, N, D, k, false, d_out_idx, d_out_dist);
}
template void knn_device_self<float>(wtp_ctx*, const float*, int64_t, int, int, int64_t*, float*);
template void knn_device_self<double>(wtp_ctx*, const double*, int64_t, int, int, int64_t*, double*);

template <class T>
static int32_t knn_host(wtp_ctx* ctx, const T* pts, int64_t N, int32_t D, int32_t k, bool drop_first, int64_t* out_idx, T* out_dist) {
    API_BEGIN(ctx)
    WTP_REQUIRE(pts && out_idx && N > 0, WTP_ERR_BAD_ARG, "null pointer or empty point set");
    WTP_REQUIRE(D == 2 || D == 3, WTP_ERR_BAD_ARG, "D must be 2 or 3");
    WTP_REQUIRE(k >= 1, WTP_ERR_BAD_ARG, "k must be >= 1");
    ctx->timer.reset(ctx->stream);
    ctx->timer.begin_total();
    T* d_pts = ctx->d_pts.as<T>((size_t)N * D);
    // The ranks of a communicator share one host: uploading the whole point set on every rank moves it G times through
    // the host's memory system (5.2 ms for 8 x 120 MB on the 8-GPU box against 2.2 ms for one copy). Every rank uploads
    // the points of its own caller range instead and the ranges are all-gathered over NVLink.
    const bool sliced = ctx->world > 1 && ctx->nccl_comm && N >= (int64_t)ctx->world * 4096 && std::getenv("WTP_NO_SLICED_H2D") == nullptr;
    int64_t h2d_bytes = N * (int64_t)D * (int64_t)sizeof(T);
    {
        ScopedPhase ph(ctx->timer, PH_H2D);
        if (sliced) {
            const int64_t b = wtp_shard_begin(N, ctx->rank, ctx->world), e = wtp_shard_end(N, ctx->rank, ctx->world);
            h2d_bytes = (e - b) * (int64_t)D * (int64_t)sizeof(T);
            WTP_CUDA_CHECK(cudaMemcpyAsync(d_pts + b * D, pts + b * D, (size_t)h2d_bytes, cudaMemcpyHostToDevice, ctx->stream));
            comm_allgather_rows(ctx, d_pts, N, (size_t)D * sizeof(T));
        } else {
            WTP_CUDA_CHECK(cudaMemcpyAsync(d_pts, pts, (size_t)N * D * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    knn_device<T>(ctx, d_pts, N, D, k, drop_first, nullptr, nullptr, out_idx, out_dist);
    ctx->last_timing.bytes_h2d = h2d_bytes;
    wtp_timing keep = ctx->last_timing;
    ctx->timer.end_total();
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->last_timing = keep;
    API_END(ctx)
}

template <class T>
static int32_t knn_dev(wtp_ctx* ctx, const T* d_pts, int64_t N, int32_t D, int32_t k, void* d_out_idx, T* d_out_dist, bool out32 = false) {
    API_BEGIN(ctx)
    WTP_REQUIRE(d_pts && d_out_idx && N > 0, WTP_ERR_BAD_ARG, "null pointer or empty point set");
    ctx->timer.reset(ctx->stream);
    ctx->timer.begin_total();
    knn_device<T>(ctx, d_pts, N, D, k, true, static_cast<int64_t*>(d_out_idx), d_out_dist, nullptr, nullptr, out32);
    ctx->timer.end_total();
    API_END(ctx)
}

}  // namespace wtp

namespace wtp {
// On a multi-device context every child answers its shard of the same call (one host thread per device): with the row
// exchange each fills one contiguous part of the caller's table, together the whole table. Small point sets (fewer than
// 4096 per device) are not worth sharding and run on the single-device context.
template <class T>
static int32_t knn_entry(wtp_ctx* c, const T* p, int64_t N, int32_t D, int32_t k, bool drop, int64_t* oi, T* od) {
    if (is_multi(c) && N >= (int64_t)4096 * (int64_t)c->children.size())
        return multi_run(c, [&](wtp_ctx* ch, int) { return knn_host<T>(ch, p, N, D, k, drop, oi, od); });
    return knn_host<T>(solo_of(c), p, N, D, k, drop, oi, od);
}
static int32_t no_device_pointers(wtp_ctx* c) {
    return fail(c, Error{WTP_ERR_UNSUPPORTED, "device-pointer entry points belong to one device: use a single-device context"});
}
}  // namespace wtp

extern "C" {

int32_t wtp_knn_f32(wtp_ctx* c, const float* p, int64_t N, int32_t D, int32_t k, int64_t* oi, float* od) { return knn_entry<float>(c, p, N, D, k, true, oi, od); }
int32_t wtp_knn_f64(wtp_ctx* c, const double* p, int64_t N, int32_t D, int32_t k, int64_t* oi, double* od) { return knn_entry<double>(c, p, N, D, k, true, oi, od); }
int32_t wtp_knn_self_f32(wtp_ctx* c, const float* p, int64_t N, int32_t D, int32_t k, int64_t* oi, float* od) { return knn_entry<float>(c, p, N, D, k, false, oi, od); }
int32_t wtp_knn_self_f64(wtp_ctx* c, const double* p, int64_t N, int32_t D, int32_t k, int64_t* oi, double* od) { return knn_entry<double>(c, p, N, D, k, false, oi, od); }
int32_t wtp_knn_dev_f32(wtp_ctx* c, const float* p, int64_t N, int32_t D, int32_t k, int64_t* oi, float* od) { return is_multi(c) ? no_device_pointers(c) : knn_dev<float>(c, p, N, D, k, oi, od); }
int32_t wtp_knn_dev_f64(wtp_ctx* c, const double* p, int64_t N, int32_t D, int32_t k, int64_t* oi, double* od) { return is_multi(c) ? no_device_pointers(c) : knn_dev<double>(c, p, N, D, k, oi, od); }
int32_t wtp_knn_dev_i32_f32(wtp_ctx* c, const float* p, int64_t N, int32_t D, int32_t k, int32_t* oi, float* od) { return is_multi(c) ? no_device_pointers(c) : knn_dev<float>(c, p, N, D, k, oi, od, true); }
int32_t wtp_knn_dev_i32_f64(wtp_ctx* c, const double* p, int64_t N, int32_t D, int32_t k, int32_t* oi, double* od) { return is_multi(c) ? no_device_pointers(c) : knn_dev<double>(c, p, N, D, k, oi, od, true); }

int64_t wtp_shard_owned_count(const wtp_ctx* ctx) { return ctx ? ctx->owned_end - ctx->owned_begin : -1; }
int32_t wtp_shard_owned_dev(wtp_ctx* ctx, int64_t* d_ids) {
    if (is_multi(ctx)) return no_device_pointers(ctx);
    API_BEGIN(ctx)
    WTP_REQUIRE(d_ids || ctx->owned_end == ctx->owned_begin, WTP_ERR_BAD_ARG, "null output");
    if (ctx->owned_contiguous) {   // the row exchange: the caller range [owned_begin, owned_end) itself
        std::vector<int64_t> ids((size_t)(ctx->owned_end - ctx->owned_begin));
        for (size_t t = 0; t < ids.size(); ++t) ids[t] = ctx->owned_begin + (int64_t)t + 1;
        WTP_CUDA_CHECK(cudaMemcpyAsync(d_ids, ids.data(), ids.size() * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    } else {
        owned_ids(ctx, ctx->index[0], ctx->owned_f64, ctx->owned_begin, ctx->owned_end, d_ids);
    }
    API_END(ctx)
}
int32_t wtp_shard_owned(wtp_ctx* ctx, int64_t* ids) {
    API_BEGIN(ctx)
    const int64_t n = ctx->owned_end - ctx->owned_begin;
    WTP_REQUIRE(ids || n == 0, WTP_ERR_BAD_ARG, "null output");
    if (ctx->owned_contiguous) {
        for (int64_t t = 0; t < n; ++t) ids[t] = ctx->owned_begin + t + 1;
    } else if (n > 0) {
        int64_t* d_ids = ctx->d_misc.as<int64_t>((size_t)n);
        owned_ids(ctx, ctx->index[0], ctx->owned_f64, ctx->owned_begin, ctx->owned_end, d_ids);
        WTP_CUDA_CHECK(cudaMemcpyAsync(ids, d_ids, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    API_END(ctx)
}

}  // extern "C"
