// multi.h — dispatch of host entry points on a single-process multi-device context (wtp_create_multi, comm.cu).
#pragma once
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace wtp {

// The in-process side of a multi-device context: a reusable barrier and a table the ranks use to hand each other device
// pointers (the peers' buffers are reachable directly once peer access is enabled: no IPC handles inside one process).
struct LocalGroup {
    int world = 0;
    std::mutex m;
    std::condition_variable cv;
    int waiting = 0;
    uint64_t generation = 0;
    void* ptrs[2][WTP_MAX_PEERS] = {};   // [which PeerSet][rank]
    int flags[WTP_MAX_PEERS] = {};
    void barrier() {
        std::unique_lock<std::mutex> lk(m);
        const uint64_t gen = generation;
        if (++waiting == world) { waiting = 0; ++generation; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != gen; });
    }
};

inline bool is_multi(const wtp_ctx* c) { return c && !c->children.empty(); }
// entry points that do not shard run on the parent's single-device context
inline wtp_ctx* solo_of(wtp_ctx* c) { return is_multi(c) ? c->solo : c; }

uint64_t next_error_stamp();   // api.cu

// f(child, rank) on every child at once, one host thread per device. Returns the status of the lowest failing rank (0 if
// none) and leaves its message on the parent.
template <class F>
int32_t multi_run(wtp_ctx* parent, F&& f) {
    const int G = (int)parent->children.size();
    std::vector<int32_t> rc((size_t)G, 0);
    std::vector<std::thread> th;
    th.reserve((size_t)G);
    for (int r = 1; r < G; ++r) th.emplace_back([&, r] { rc[(size_t)r] = f(parent->children[(size_t)r], r); });
    rc[0] = f(parent->children[0], 0);
    for (auto& t : th) t.join();
    for (int r = 0; r < G; ++r)
        if (rc[(size_t)r] != 0) {
            parent->last_error = "device " + std::to_string(parent->children[(size_t)r]->device) + ": " + parent->children[(size_t)r]->last_error;
            parent->error_stamp = next_error_stamp();
            return rc[(size_t)r];
        }
    return 0;
}

}  // namespace wtp
