// common.cuh — shared declarations of libwtp_cuda.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/wtp_cuda.h"

namespace wtp {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this
constexpr int kWarp = 32;
}
constexpr int WTP_MAX_PEERS = 8;  // ranks of one NVSwitch domain whose buffers a context can map
namespace wtp {

// ------------------------------------------------------------------ errors
struct Error {
    int32_t status;
    std::string msg;
};

#define WTP_CUDA_CHECK(expr)                                                                     \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            throw ::wtp::Error{_e == cudaErrorMemoryAllocation ? WTP_ERR_OOM : WTP_ERR_CUDA,     \
                               std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " +     \
                                   __FILE__ + ":" + std::to_string(__LINE__)};                   \
        }                                                                                        \
    } while (0)

#define WTP_REQUIRE(cond, status, message)                     \
    do {                                                       \
        if (!(cond)) throw ::wtp::Error{(status), (message)};  \
    } while (0)

// --------------------------------------------------------- device records
// One sorted point: coordinates + the caller's index, 16 B (f32) / 32 B (f64), so that
// any cell range is 16-byte aligned for cp.async.bulk and one LDG.128/LDS.128 per point.
template <class T>
struct alignas(16) P4 {
    T x, y, z;
    T w;  // bit pattern of the original (0-based) index, see idx_of / set_idx
};

__host__ __device__ inline uint32_t idx_of(const P4<float>& p) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(p.w);
#else
    uint32_t u; memcpy(&u, &p.w, 4); return u;
#endif
}
__host__ __device__ inline uint32_t idx_of(const P4<double>& p) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__double_as_longlong(p.w);
#else
    uint64_t u; memcpy(&u, &p.w, 8); return (uint32_t)u;
#endif
}
__device__ inline float idx_bits(float, uint32_t i) { return __uint_as_float(i); }
__device__ inline double idx_bits(double, uint32_t i) { return __longlong_as_double((long long)i); }

// Uniform grid over the bounding box; row-major cells (x fastest) so that the cells of
// one x-row are contiguous in the sorted array.
template <class T>
struct Grid {
    T lo[3];
    T inv_c;     // 1 / cell size
    T c;         // cell size
    T slack;     // absolute safety margin for pruning bounds (few ulps of the extent)
    int n[3];    // cells per dimension (n[2] = 1 in 2-D)
    uint32_t ncells;
    // Layers of the slowest axis (z in 3-D, y in 2-D) that the index holds, inclusive. The whole grid, except for
    // the windowed index of a sharded k-NN call (grid.cu, build_index_window): a search that needs a layer outside
    // the window must say so instead of reading it.
    int w_lo, w_hi;
};

template <class T>
__host__ __device__ inline int cell_coord(const Grid<T>& g, T v, int d) {
    // monotone in v: floor(fl(fl(v - lo) * inv_c)), clamped (points outside the box of a
    // stale grid land in the border cells, which are treated as unbounded outward).
    T t = (v - g.lo[d]) * g.inv_c;
    int i = (int)floor(t);
    i = i < 0 ? 0 : i;
    i = i > g.n[d] - 1 ? g.n[d] - 1 : i;
    return i;
}

// Morton bit spreading (bvh.cu, mesh.cu)
__device__ __forceinline__ uint32_t spread3(uint32_t v) {  // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ uint32_t spread2(uint32_t v) {  // 16 bits -> every second bit
    v &= 0xffffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

// ------------------------------------------------- the library's random stream
// _safe_direction (src/repel.jl:358-364) draws a random unit vector for a coincident neighbour from Julia's global
// RNG; that stream cannot be reproduced outside Julia, so the library defines its own, counter-based one (documented
// in include/wtp_cuda.h, restated in the oracle): the direction for (sweep key, point i, neighbour j) is the first
// point of a hashed sequence in the cube [-1, 1)^D that falls inside the unit ball (and not within 2^-5 of the
// origin), normalised — uniform on the circle / sphere like randn/norm. Only exactly rounded operations (integer
// hash, int -> T conversion, multiply, add, sqrt, divide, none fused), so every implementation gives the same bits.
__host__ __device__ inline uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
// key of one sweep: the seed of the call and the 1-based iteration number
__host__ __device__ inline uint64_t sweep_key(uint64_t seed, uint64_t iteration) { return mix64(seed ^ mix64(iteration)); }

// ------------------------------------------------------------ device memory
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    void* reserve(size_t bytes) {
        if (bytes > cap) {
            if (p) { cudaFree(p); p = nullptr; cap = 0; }
            size_t want = bytes + bytes / 8 + 256;
            cudaError_t e = cudaMalloc(&p, want);
            if (e != cudaSuccess) {
                p = nullptr; cap = 0;
                (void)cudaGetLastError();
                throw Error{WTP_ERR_OOM, "cudaMalloc(" + std::to_string(want) + " bytes): " + cudaGetErrorString(e)};
            }
            cap = want;
        }
        return p;
    }
    template <class U> U* as(size_t count) { return static_cast<U*>(reserve(count * sizeof(U))); }
    template <class U> U* get() const { return static_cast<U*>(p); }
};

// --------------------------------------------------------------- timing
enum Phase { PH_H2D = 0, PH_BBOX, PH_CELLKEY, PH_SORT, PH_REORDER, PH_QUERY, PH_SCAN, PH_REDUCE, PH_COMM, PH_D2H, PH_COUNT };

struct PhaseTimer {
    bool enabled = false;
    cudaStream_t stream = nullptr;
    struct Span { int phase; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    cudaEvent_t total_a = nullptr, total_b = nullptr;
    bool have_total = false;
    cudaEvent_t get() {
        if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        return pool[used++];
    }
    void reset(cudaStream_t s) { stream = s; spans.clear(); used = 0; have_total = false; }
    void begin_total() { if (enabled) { total_a = get(); cudaEventRecord(total_a, stream); } }
    void end_total() { if (enabled) { total_b = get(); cudaEventRecord(total_b, stream); have_total = true; } }
    int open(int phase) {
        if (!enabled) return -1;
        Span s{phase, get(), nullptr};
        cudaEventRecord(s.a, stream);
        spans.push_back(s);
        return (int)spans.size() - 1;
    }
    void close(int h) {
        if (h < 0) return;
        spans[h].b = get();
        cudaEventRecord(spans[h].b, stream);
    }
    ~PhaseTimer() { for (auto e : pool) cudaEventDestroy(e); }
};

struct ScopedPhase {
    PhaseTimer& t; int h;
    ScopedPhase(PhaseTimer& t_, int phase) : t(t_), h(t_.open(phase)) {}
    ~ScopedPhase() { t.close(h); }
};

// -------------------------------------------------------------- the index
// Device-resident spatial index of one point set (all buffers owned by the context).
struct IndexBuffers {
    DevBuf keys_a, keys_b, vals_a, vals_b;  // radix sort ping-pong (u32)
    DevBuf block_hist;                      // 256 x numBlocks digit histogram
    DevBuf scan_tmp;                        // block sums of the scans
    DevBuf sorted;                          // P4<T>[N]
    DevBuf staging;                         // counting build: caller indices grouped by cell, arrival order inside a cell
    DevBuf cell_start;                      // u32[ncells + 1]
    DevBuf bbox_partial;                    // per-block min/max
    DevBuf bbox;                            // 6 x T
    DevBuf unit_hist;                       // windowed build: points per layer of the slowest axis + the picked window
    int64_t cs_rebase = 0;                  // windowed build: cell_start[0] belongs to this cell id (0: whole grid)
    const uint32_t* cells() const { return cell_start.get<uint32_t>() - cs_rebase; }   // indexable by global cell id
};

// The part of the sorted order a windowed index holds: the points of the cell ids [key_lo, key_hi), which are the
// sorted positions [P0, P0 + M) of the whole set.
struct IndexWindow {
    int64_t P0 = 0, M = 0;
    uint32_t key_lo = 0, key_hi = 0;
    int w_lo = 0, w_hi = 0;
};

// Implicit BVH over the Morton-sorted boundary set of a variable spacing (bvh.cu).
struct BvhBuffers {
    IndexBuffers ib;   // sort scratch + sorted P4 records
    DevBuf boxes;      // Box<T>[2 * leaf_pow2], heap layout, node 1 = root
    int64_t n = 0;
    int64_t leaf_pow2 = 0;
};

// Triangle mesh of the wall rule (mesh.cu): sorted triangle records, implicit BVH, the
// per-triangle feature pseudonormals in caller order, and the per-point wall state.
struct MeshBuffers {
    IndexBuffers ib;                   // sort scratch
    DevBuf raw;                        // n_tri x 9 of T as uploaded
    DevBuf tris;                       // TriRec<T>[n_tri], Morton order
    DevBuf boxes;                      // Box<T>[2 * leaf_pow2]
    DevBuf fnorm;                      // n_tri x 21 of T, caller order
    DevBuf is_bnd, tri_idx, escaped;   // u8 / i64 / u8 per movable point
    DevBuf hint;                       // u32 per movable point: sorted position of last nearest triangle
    int64_t n = 0, leaf_pow2 = 0;
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    double offset = 0, maxabs = 0;
};

// Where the row of a query goes in the output table: at its caller index (minus q_begin), or, for the compact
// table of a sharded host call, at its sorted position (minus pos_base).
// Row exchange of a sharded host call (n_owners > 0): the row of caller index i belongs to the rank that owns the caller
// range holding i and is stored straight into that rank's buffer (peer memory over NVLink), so that every rank ends up
// with the rows of one contiguous caller range: the element address is out + owner_delta[r] + i * k_out (owner_delta in
// table elements; the peers' buffers are mapped into this process's address space).
struct RowMap {
    uint32_t q_begin, pos_base;
    int by_position;
    int n_owners;
    uint32_t owner_begin[WTP_MAX_PEERS + 1];
    long long owner_delta[WTP_MAX_PEERS];
    __host__ __device__ int64_t row(uint32_t sorted_pos, uint32_t caller_idx) const {
        return by_position ? (int64_t)(sorted_pos - pos_base) : (int64_t)(caller_idx - q_begin);
    }
    // offset of the row's first element in the output table
    __host__ __device__ int64_t elem(uint32_t sorted_pos, uint32_t caller_idx, int k_out) const {
        if (n_owners > 0) {
            long long delta = owner_delta[0];          // constant indices only (the loop unrolls): the arrays stay in the parameter bank
            for (int t = 1; t < WTP_MAX_PEERS; ++t) delta = (t < n_owners && caller_idx >= owner_begin[t]) ? owner_delta[t] : delta;
            return (int64_t)delta + (int64_t)caller_idx * k_out;
        }
        return row(sorted_pos, caller_idx) * k_out;
    }
};

// Why a query left the fast path of a tiled k-NN pass (knn_tile.cuh). TK_SPARSE: too few candidates in its block,
// or the K-th neighbour is not provably inside it (the grid is too fine there). TK_DENSE: the slab does not fit the
// tile, or more than TK_LCAP hits (too coarse there). TK_OTHER: look-alike key images or exact ties. Anything but
// TK_OK is appended to one list of sorted positions and answered by the general kernel; the per-kind counters are
// statistics.
enum { TK_OK = 0, TK_SPARSE = 1, TK_DENSE = 2, TK_OTHER = 3, TK_NOFIT = 4 };   // TK_NOFIT: the slab does not fit the tile (counted apart, reported with the dense ones)
struct TileFails {
    uint32_t* counters;    // [0] length of the list, [1] sparse, [2] dense, [3] other
    uint32_t* list;        // sorted positions
};

// Buffers that every rank of a communicator maps from every other rank (CUDA IPC over NVLink; comm.cu, comm_peer_buffers):
// two per rank, alternating per use, so that a rank running ahead never writes into a buffer a peer is still reading.
struct PeerSet {
    DevBuf own;                      // this rank's buffers (2 x bytes_each)
    size_t bytes_each = 0;
    void* base[WTP_MAX_PEERS] = {};  // base[r]: rank r's buffers in this process's address space (own for r == rank)
    bool mapped = false, unavailable = false;
};

struct NcclApi;  // comm.cu
class HostPool;  // host_pool.h
struct LocalGroup;  // comm.cu: the ranks of a single-process multi-device context (wtp_create_multi)

}  // namespace wtp

// The opaque context of the C ABI.
struct wtp_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // D2H of finished chunks while the next chunk computes
    cudaEvent_t ev_chunk[8] = {};
    cudaEvent_t ev_copy_done = nullptr;
    std::string last_error;
    double cell_occupancy = 0.0;  // <= 0: default per dimension
    int64_t launches = 0;
    wtp::PhaseTimer timer;
    wtp_timing last_timing{};
    // spatial index of the queried point set / the repel snapshot
    wtp::IndexBuffers index[3];          // [0]: every call; [1], [2]: the coarser density classes of a graded repel (repel.cu)
    // 1-NN structure of a variable spacing's boundary set
    wtp::BvhBuffers bvh;
    // triangle mesh of the wall rule (3-argument repel) and of the batched mesh queries
    wtp::MeshBuffers mesh;
    wtp::DevBuf d_pts, d_out_idx, d_out_dist, d_offsets, d_counts, d_indices, d_misc, d_misc2;
    wtp::DevBuf d_spacing_pts, d_spacings, d_p_new, d_reduce, d_qlist, d_nn, d_fail, d_pack;
    size_t last_d2h_bytes = 0;           // bytes the last widening copy moved over PCIe
    // radius two-call state
    struct {
        bool pending = false; bool f64 = false; bool dev_input = false;
        int64_t N = 0; int D = 0; double r = 0; int64_t nnz = -1;
        const void* pts = nullptr;
        int64_t q_begin = 0, q_end = 0;
    } radius;
    unsigned char grid_storage[2][128];  // last Grid<T> per index (host copy)
    int64_t owned_begin = 0, owned_end = 0; bool owned_f64 = false;   // sorted range answered by the last sharded k-NN call
    bool window_off = false;     // a windowed k-NN call had to fall back (graded cloud): later calls build the whole index at once
    int64_t last_window_points = 0, last_window_missed = 0;   // statistics of the last sharded k-NN call
    int64_t last_tile_sparse = 0, last_tile_dense = 0, last_tile_other = 0;   // leftovers of the last tiled sweep, by kind
    // multi-GPU
    int rank = 0, world = 1;
    void* nccl_comm = nullptr;
    wtp::NcclApi* nccl = nullptr;
    // Single-process multi-device context (wtp_create_multi, comm.cu). The handle the caller holds is a PARENT: it owns
    // one child context per device (ranks of one NCCL communicator made with ncclCommInitAll, sharing a LocalGroup for
    // the in-process pointer exchange and barriers) and a plain single-device context (`solo`, on the first device) for
    // the entry points that do not shard. A host entry point called on the parent runs on every child at once, one
    // host thread per device; the children answer their shards exactly as one-process-per-GPU ranks would.
    std::vector<wtp_ctx*> children;
    std::vector<int64_t> children_base;  // parent, radius two-call state: where each child's entries start in the caller's CSR
    wtp_ctx* solo = nullptr;
    wtp::LocalGroup* group = nullptr;    // children: the group they belong to (owned by the parent)
    bool quiet = false;                  // children of rank > 0 during a replicated call (repel): results are written back by rank 0 only
    uint64_t error_stamp = 0;            // order of the last failure among the contexts of one parent
    // peer-memory exchange of the run-sharded repel (comm.cu, comm_peer_buffers): every rank's two run buffers, mapped
    // into this process with CUDA IPC, so that the sweep kernels store their records straight into all ranks' buffers
    // over NVLink instead of an all-gather afterwards
    wtp::PeerSet peers;                  // run buffers of the run-sharded repel
    // row buffers of the sharded host k-NN: every rank's sweep stores the row of caller index i into the buffer of the rank
    // that owns the caller range of i, so each rank brings back one contiguous part of the caller's table
    wtp::PeerSet row_peers;
    uint64_t row_exchanges = 0;          // calls that used the row exchange (buffer parity)
    bool last_d2h_direct = false;         // ... and its rows went back as int64 written by the DMA engine (pinned caller table)
    bool owned_contiguous = false;       // the last sharded k-NN call answered a contiguous caller range [owned_begin, owned_end)
    // host k-NN pipeline: 4-byte index staging (pinned ring) and the widening threads
    void* h_stage = nullptr;
    size_t h_stage_slot_bytes = 0;
    cudaEvent_t ev_copied[4] = {};
    std::vector<cudaEvent_t> ev_ring;     // one event per slot of the staging ring (d2h_pipeline.h)
    int h_stage_ring = 0;                  // slots in the ring
    wtp::HostPool* pool = nullptr;
    void* h_ids = nullptr;            // pinned, grow-only: caller indices of a sharded host call's rows (4 bytes each)
    size_t h_ids_bytes = 0;
    // pinned staging for scalar read-backs
    void* h_pinned = nullptr;
    size_t h_pinned_bytes = 0;
};
