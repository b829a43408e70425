// comm.cu — multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// NCCL is bound at run time (dlopen "libnccl.so.2") so that the library shares the NCCL
// the host program already loaded (torch's bundled copy under torchrun) and needs no
// link-time dependency for single-GPU use. Collectives on the data path:
//   k-NN / radius : none (queries shard by contiguous range, index replicated)
//   repel         : per iteration the ranks' runs of moved points meet in every rank's buffer — stored there by the
//                   sweep kernels themselves through peer memory (comm_peer_buffers: CUDA IPC over NVLink), or by one
//                   all-gather after the sweep when the ranks cannot map each other — + one all-gather of the per-rank
//                   stop-test partials (about a hundred bytes), which is also the barrier of the peer-memory path
#include <dlfcn.h>

#include <vector>

#include "kernels.cuh"
#include "multi.h"

namespace wtp {

typedef struct { char internal[128]; } nccl_unique_id;
typedef void* nccl_comm_t;
enum { NCCL_INT8 = 0 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(nccl_unique_id*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
    int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

static NcclApi* load_nccl() {
    static NcclApi api;
    if (api.handle) return &api;
    const char* env = getenv("WTP_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) throw Error{WTP_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found")};
#define BIND(field, sym)                                                                     \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, sym));                        \
    if (!api.field) throw Error{WTP_ERR_NCCL, std::string("libnccl is missing symbol ") + sym};
    BIND(GetUniqueId, "ncclGetUniqueId")
    BIND(CommInitRank, "ncclCommInitRank")
    BIND(CommInitAll, "ncclCommInitAll")
    BIND(CommDestroy, "ncclCommDestroy")
    BIND(AllGather, "ncclAllGather")
    BIND(Broadcast, "ncclBroadcast")
    BIND(GroupStart, "ncclGroupStart")
    BIND(GroupEnd, "ncclGroupEnd")
    BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
    api.handle = h;
    return &api;
}

#define NCCL_CHECK(api, expr)                                                                          \
    do {                                                                                               \
        int _r = (expr);                                                                               \
        if (_r != 0) throw Error{WTP_ERR_NCCL, std::string(#expr) + ": " + (api)->GetErrorString(_r)}; \
    } while (0)

// every rank owns rows [shard_begin, shard_end) of d_buf; afterwards all ranks hold all rows
void comm_allgather_rows(wtp_ctx* ctx, void* d_buf, int64_t n_rows, size_t row_bytes) {
    if (ctx->world <= 1) return;
    NcclApi* api = ctx->nccl;
    NCCL_CHECK(api, api->GroupStart());
    for (int r = 0; r < ctx->world; ++r) {
        const int64_t b = wtp_shard_begin(n_rows, r, ctx->world), e = wtp_shard_end(n_rows, r, ctx->world);
        if (e <= b) continue;
        char* p = static_cast<char*>(d_buf) + (size_t)b * row_bytes;
        NCCL_CHECK(api, api->Broadcast(p, p, (size_t)(e - b) * row_bytes, NCCL_INT8, r, ctx->nccl_comm, ctx->stream));
    }
    NCCL_CHECK(api, api->GroupEnd());
}

void comm_allgather_fixed(wtp_ctx* ctx, const void* d_in, void* d_out, size_t bytes_per_rank) {
    if (ctx->world <= 1) {
        WTP_CUDA_CHECK(cudaMemcpyAsync(d_out, d_in, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
        return;
    }
    NcclApi* api = ctx->nccl;
    NCCL_CHECK(api, api->AllGather(d_in, d_out, bytes_per_rank, NCCL_INT8, ctx->nccl_comm, ctx->stream));
}

// Peer buffers of the run-sharded repel: every rank allocates 2 x bytes_each, the CUDA IPC handles go round with one
// small ncclAllGather, and every rank maps the others' buffers (NVLink peer access). Returns false — and the caller
// keeps the NCCL all-gather — when the ranks are not all peers of each other or IPC is not available (decided
// collectively, so every rank takes the same path). Collective: every rank must call it with the same size.
static void peers_unmap(wtp_ctx* ctx, PeerSet& pb) {
    for (int r = 0; r < WTP_MAX_PEERS; ++r) {
        if (pb.base[r] && r != ctx->rank && !ctx->group) cudaIpcCloseMemHandle(pb.base[r]);   // in-process peers: plain pointers
        pb.base[r] = nullptr;
    }
    pb.mapped = false;
}

// The ranks of a single-process multi-device context reach each other's buffers directly (peer access was enabled when
// the context was made): the pointers go round through the group's table, two in-process barriers instead of handles.
static bool group_peer_buffers(wtp_ctx* ctx, PeerSet& pb, size_t bytes_each) {
    LocalGroup* g = ctx->group;
    const int which = &pb == &ctx->peers ? 0 : 1;
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    g->barrier();                                           // nobody is still using the old buffers
    if (pb.own.p) { cudaFree(pb.own.p); pb.own.p = nullptr; pb.own.cap = 0; }
    const size_t each = (bytes_each + bytes_each / 8 + 4095) & ~(size_t)4095;
    void* mine = nullptr;
    const bool ok = cudaMalloc(&mine, 2 * each) == cudaSuccess;
    (void)cudaGetLastError();
    g->ptrs[which][ctx->rank] = ok ? mine : nullptr;
    g->barrier();
    bool all_ok = true;
    for (int r = 0; r < ctx->world; ++r) all_ok = all_ok && g->ptrs[which][r] != nullptr;
    for (int r = 0; r < ctx->world; ++r) pb.base[r] = all_ok ? g->ptrs[which][r] : nullptr;
    g->barrier();                                           // everyone has read the table
    pb.own.p = mine; pb.own.cap = mine ? 2 * each : 0;
    pb.bytes_each = each;
    pb.mapped = all_ok;
    pb.unavailable = !all_ok;
    return all_ok;
}

bool comm_peer_buffers(wtp_ctx* ctx, PeerSet& pb, size_t bytes_each) {
    if (ctx->world <= 1 || ctx->world > WTP_MAX_PEERS || !ctx->nccl_comm || pb.unavailable || getenv("WTP_NO_P2P")) return false;
    if (pb.mapped && pb.bytes_each >= bytes_each) return true;
    if (ctx->group) return group_peer_buffers(ctx, pb, bytes_each);
    NcclApi* api = ctx->nccl;
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (pb.mapped) {
        // growing: every rank closes its mappings first, and only after all have done so (a tiny all-gather as the
        // barrier) does anybody free the buffer the others had mapped
        peers_unmap(ctx, pb);
        int64_t* d_flag = ctx->d_misc.as<int64_t>((size_t)ctx->world + 1);
        NCCL_CHECK(api, api->AllGather(d_flag + ctx->world, d_flag, sizeof(int64_t), NCCL_INT8, ctx->nccl_comm, ctx->stream));
        WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    if (pb.own.p) { cudaFree(pb.own.p); pb.own.p = nullptr; pb.own.cap = 0; }     // a fresh allocation: the handle names the whole allocation
    const size_t each = (bytes_each + bytes_each / 8 + 4095) & ~(size_t)4095;
    void* mine = nullptr;
    cudaIpcMemHandle_t handle;
    int ok = cudaMalloc(&mine, 2 * each) == cudaSuccess && cudaIpcGetMemHandle(&handle, mine) == cudaSuccess ? 1 : 0;
    (void)cudaGetLastError();
    // all-gather {ok, handle} (72 bytes per rank)
    struct Msg { int64_t ok; cudaIpcMemHandle_t h; };
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    Msg m{ok, {}};
    if (ok) m.h = handle;
    Msg* d_msg = ctx->d_misc.as<Msg>((size_t)ctx->world + 1);
    std::vector<Msg> all((size_t)ctx->world);
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_msg + ctx->world, &m, sizeof(Msg), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_CHECK(api, api->AllGather(d_msg + ctx->world, d_msg, sizeof(Msg), NCCL_INT8, ctx->nccl_comm, ctx->stream));
    WTP_CUDA_CHECK(cudaMemcpyAsync(all.data(), d_msg, sizeof(Msg) * ctx->world, cudaMemcpyDeviceToHost, ctx->stream));
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    int all_ok = 1;
    for (int r = 0; r < ctx->world; ++r) all_ok &= all[(size_t)r].ok ? 1 : 0;
    if (all_ok) {
        for (int r = 0; r < ctx->world && all_ok; ++r) {
            if (r == ctx->rank) { pb.base[r] = mine; continue; }
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[(size_t)r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); all_ok = 0; }
            else pb.base[r] = p;
        }
    }
    // second round: did every rank map every peer?
    m.ok = all_ok;
    WTP_CUDA_CHECK(cudaMemcpyAsync(d_msg + ctx->world, &m, sizeof(Msg), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_CHECK(api, api->AllGather(d_msg + ctx->world, d_msg, sizeof(Msg), NCCL_INT8, ctx->nccl_comm, ctx->stream));
    WTP_CUDA_CHECK(cudaMemcpyAsync(all.data(), d_msg, sizeof(Msg) * ctx->world, cudaMemcpyDeviceToHost, ctx->stream));
    WTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < ctx->world; ++r) all_ok &= all[(size_t)r].ok ? 1 : 0;
    if (!all_ok) {
        pb.base[ctx->rank] = nullptr;
        peers_unmap(ctx, pb);
        pb.own.p = mine; pb.own.cap = mine ? 2 * each : 0;   // freed with the context, when no peer can still have it mapped
        pb.unavailable = true;
        return false;
    }
    pb.own.p = mine; pb.own.cap = 2 * each;
    pb.bytes_each = each;
    pb.mapped = true;
    return true;
}

int32_t fail(wtp_ctx* ctx, const Error& e);

}  // namespace wtp

using namespace wtp;

extern "C" {

void wtp_comm_destroy_internal(wtp_ctx* ctx) {
    wtp::peers_unmap(ctx, ctx->peers);
    wtp::peers_unmap(ctx, ctx->row_peers);
    if (ctx->nccl_comm && ctx->nccl) ctx->nccl->CommDestroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->rank = 0;
    ctx->world = 1;
}

int32_t wtp_comm_unique_id(void* out128) {
    if (!out128) return WTP_ERR_BAD_ARG;
    try {
        NcclApi* api = load_nccl();
        nccl_unique_id id;
        int r = api->GetUniqueId(&id);
        if (r != 0) return WTP_ERR_NCCL;
        memcpy(out128, &id, sizeof(id));
    } catch (const Error&) {
        return WTP_ERR_NCCL;
    }
    return WTP_OK;
}

int32_t wtp_comm_init(wtp_ctx* ctx, int32_t rank, int32_t world, const void* unique_id128) {
    if (!ctx) return WTP_ERR_BAD_ARG;
    try {
        WTP_REQUIRE(world >= 1 && rank >= 0 && rank < world, WTP_ERR_BAD_ARG, "bad rank/world");
        WTP_CUDA_CHECK(cudaSetDevice(ctx->device));
        wtp_comm_destroy_internal(ctx);
        if (world == 1) return WTP_OK;
        if (!unique_id128) {   // shard-only context: k-NN and radius shards need no collective (repel refuses it)
            ctx->rank = rank;
            ctx->world = world;
            return WTP_OK;
        }
        NcclApi* api = load_nccl();
        nccl_unique_id id;
        memcpy(&id, unique_id128, sizeof(id));
        nccl_comm_t comm = nullptr;
        NCCL_CHECK(api, api->CommInitRank(&comm, world, id, rank));
        ctx->nccl = api;
        ctx->nccl_comm = comm;
        ctx->rank = rank;
        ctx->world = world;
    } catch (const Error& e) {
        return fail(ctx, e);
    }
    return WTP_OK;
}

int32_t wtp_comm_rank(const wtp_ctx* ctx) { return ctx ? ctx->rank : -1; }
int32_t wtp_comm_world(const wtp_ctx* ctx) { return ctx ? ctx->world : -1; }

// One context over several GPUs of the box, in one process (wtp_create(ctx**, devices*, n)): the
// parent handle owns a child context per device — the ranks of one NCCL communicator (ncclCommInitAll), with peer access
// enabled between every pair so that the exchange kernels store straight into each other's memory — and a plain
// single-device context for the entry points that do not shard. n_devices == 1 is wtp_create.
int32_t wtp_create_multi(wtp_ctx** out, const int32_t* devices, int32_t n_devices) {
    if (!out || !devices || n_devices < 1 || n_devices > WTP_MAX_PEERS) return WTP_ERR_BAD_ARG;
    *out = nullptr;
    if (n_devices == 1) return wtp_create(out, devices[0]);
    for (int a = 0; a < n_devices; ++a)
        for (int b = 0; b < a; ++b)
            if (devices[a] == devices[b]) return WTP_ERR_BAD_ARG;
    wtp_ctx* parent = new (std::nothrow) wtp_ctx();
    if (!parent) return WTP_ERR_OOM;
    parent->device = devices[0];
    int32_t rc = WTP_OK;
    try {
        NcclApi* api = load_nccl();
        for (int r = 0; r < n_devices && rc == WTP_OK; ++r) {
            wtp_ctx* c = nullptr;
            rc = wtp_create(&c, devices[r]);
            if (rc == WTP_OK) parent->children.push_back(c);
        }
        if (rc == WTP_OK) rc = wtp_create(&parent->solo, devices[0]);
        if (rc == WTP_OK) {
            for (int a = 0; a < n_devices; ++a) {              // peer access between every pair (NVLink / NVSwitch)
                WTP_CUDA_CHECK(cudaSetDevice(devices[a]));
                for (int b = 0; b < n_devices; ++b) {
                    if (a == b) continue;
                    int can = 0;
                    WTP_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, devices[a], devices[b]));
                    WTP_REQUIRE(can, WTP_ERR_CUDA, "wtp_create_multi: the devices cannot access each other's memory");
                    const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) WTP_CUDA_CHECK(e);
                    (void)cudaGetLastError();
                }
            }
            std::vector<nccl_comm_t> comms((size_t)n_devices, nullptr);
            std::vector<int> devs(devices, devices + n_devices);
            NCCL_CHECK(api, api->CommInitAll(comms.data(), n_devices, devs.data()));
            LocalGroup* g = new LocalGroup();
            g->world = n_devices;
            parent->group = g;
            for (int r = 0; r < n_devices; ++r) {
                wtp_ctx* c = parent->children[(size_t)r];
                c->nccl = api; c->nccl_comm = comms[(size_t)r]; c->rank = r; c->world = n_devices; c->group = g;
            }
        }
    } catch (const Error& e) {
        rc = e.status;
    } catch (...) {
        rc = WTP_ERR_CUDA;
    }
    if (rc != WTP_OK) { wtp_destroy(parent); return rc; }
    *out = parent;
    return WTP_OK;
}

}  // extern "C"
