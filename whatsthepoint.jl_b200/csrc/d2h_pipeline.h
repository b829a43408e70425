// d2h_pipeline.h — device -> pinned staging ring -> caller's host array, chunk by chunk, without a barrier per chunk.
//
// The host entry points bring their results back through a ring of pinned staging slots on the copy stream and let the
// host pool turn every chunk into the caller's layout (4-byte indices widened to int64, rows scattered to caller
// positions) while later chunks are on the wire. One thread of the pool is the producer: it enqueues the copy of chunk
// c as soon as every worker has finished chunk c - R (the slot is free), polls the copy events in order and publishes
// the number of chunks that have landed; the other threads are workers: each takes its own slice of every chunk as it
// lands. All hand-offs are single atomics (no mutex, no condition variable per chunk), so the chunks can be small
// enough (WTP_STAGE_MB) for the DMA-written data to still be in the last-level cache when the workers read it.
#pragma once
#include <cuda_runtime.h>
#include <emmintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <functional>
#include <memory>
#include <vector>

#include "common.cuh"
#include "host_pool.h"

namespace wtp {

struct alignas(64) PaddedCounter { std::atomic<int64_t> v{0}; };

// job(chunk, staged bytes of the chunk, offset of the chunk in the source, bytes, worker, workers)
using ChunkJob = std::function<void(int64_t, const char*, size_t, size_t, int, int)>;

// d_src: `total` bytes on the device, complete on ctx->stream. chunk_bytes: a multiple of the record size the job
// expects, at most the slot size. Returns after every chunk has been processed.
inline void d2h_pipeline(wtp_ctx* ctx, const void* d_src, size_t total, size_t chunk_bytes, size_t slot_bytes, int ring, const ChunkJob& job) {
    if (total == 0) return;
    const int64_t n_chunks = (int64_t)((total + chunk_bytes - 1) / chunk_bytes);
    while ((int)ctx->ev_ring.size() < ring) {
        cudaEvent_t e;
        WTP_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_ring.push_back(e);
    }
    WTP_CUDA_CHECK(cudaEventRecord(ctx->ev_chunk[0], ctx->stream));
    WTP_CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_chunk[0], 0));
    char* stage = static_cast<char*>(ctx->h_stage);
    const char* src = static_cast<const char*>(d_src);
    auto enqueue = [&](int64_t c) {
        const size_t off = (size_t)c * chunk_bytes, len = std::min(chunk_bytes, total - off);
        WTP_CUDA_CHECK(cudaMemcpyAsync(stage + (size_t)(c % ring) * slot_bytes, src + off, len, cudaMemcpyDeviceToHost, ctx->copy_stream));
        WTP_CUDA_CHECK(cudaEventRecord(ctx->ev_ring[(size_t)(c % ring)], ctx->copy_stream));
    };
    const int threads = ctx->pool->size();
    if (threads < 2) {   // one thread: copy ahead, process in order
        for (int64_t c = 0; c < std::min<int64_t>(ring - 1, n_chunks); ++c) enqueue(c);
        for (int64_t c = 0; c < n_chunks; ++c) {
            WTP_CUDA_CHECK(cudaEventSynchronize(ctx->ev_ring[(size_t)(c % ring)]));
            const size_t off = (size_t)c * chunk_bytes, len = std::min(chunk_bytes, total - off);
            job(c, stage + (size_t)(c % ring) * slot_bytes, off, len, 0, 1);
            if (c + ring - 1 < n_chunks) enqueue(c + ring - 1);
        }
    } else {
        const int workers = threads - 1;
        std::unique_ptr<PaddedCounter[]> finished(new PaddedCounter[(size_t)workers]);
        PaddedCounter landed;
        std::atomic<int> failed{0};
        const int device = ctx->device;
        const bool share_cores = ctx->world > 1;
        ctx->pool->run([&](int part, int) {
            if (part == 0) {   // producer
                try {
                    int64_t next_enqueue = 0, next_land = 0;
                    while (next_land < n_chunks) {
                        bool progressed = false;
                        if (next_enqueue < n_chunks) {
                            int64_t slowest = next_enqueue;
                            for (int w = 0; w < workers; ++w) slowest = std::min(slowest, finished[(size_t)w].v.load(std::memory_order_acquire));
                            if (next_enqueue - slowest < ring) { enqueue(next_enqueue++); progressed = true; }
                        }
                        if (next_land < next_enqueue) {
                            const cudaError_t q = cudaEventQuery(ctx->ev_ring[(size_t)(next_land % ring)]);
                            if (q == cudaSuccess) { landed.v.store(++next_land, std::memory_order_release); progressed = true; }
                            else if (q != cudaErrorNotReady) WTP_CUDA_CHECK(q);
                        }
                        if (!progressed) {
                            if (share_cores) std::this_thread::sleep_for(std::chrono::microseconds(15));   // leave the core to a worker
                            else _mm_pause();
                        }
                    }
                } catch (...) {
                    failed.store(1);
                    landed.v.store(n_chunks, std::memory_order_release);   // release the workers
                }
            } else {
                cudaSetDevice(device);
                const int w = part - 1;
                for (int64_t c = 0; c < n_chunks; ++c) {
                    while (landed.v.load(std::memory_order_acquire) <= c) _mm_pause();
                    if (!failed.load(std::memory_order_relaxed)) {
                        const size_t off = (size_t)c * chunk_bytes, len = std::min(chunk_bytes, total - off);
                        job(c, stage + (size_t)(c % ring) * slot_bytes, off, len, w, workers);
                    }
                    finished[(size_t)w].v.store(c + 1, std::memory_order_release);
                }
            }
        });
        WTP_REQUIRE(!failed.load(), WTP_ERR_CUDA, "device-to-host pipeline: a copy failed");
    }
    WTP_CUDA_CHECK(cudaEventRecord(ctx->ev_copy_done, ctx->copy_stream));
    WTP_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_done, 0));
}

}  // namespace wtp
