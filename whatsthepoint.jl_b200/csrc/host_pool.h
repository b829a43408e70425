// host_pool.h — a small persistent worker pool for the host side of the C ABI.
//
// The ABI returns int64 indices (Julia's Int), which makes the k-NN result PCIe-bound:
// 8 B per neighbour leave the device. The host entry points instead bring 4-byte indices back
// into pinned staging and widen them into the caller's int64 buffer on these threads, chunk
// by chunk, while the next chunk is still on the wire. Purely data movement: no neighbour is
// computed on the CPU.
#pragma once
#include <emmintrin.h>
#include <tmmintrin.h>
#include <stdint.h>

#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace wtp {

class HostPool {
  public:
    explicit HostPool(int n) : n_(n < 1 ? 1 : n) {
        for (int i = 1; i < n_; ++i) workers_.emplace_back([this, i] { loop(i); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            ++gen_;
        }
        cv_work_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return n_; }
    // runs f(part, parts) for part = 0..parts-1 (parts = pool size) and returns when all are done;
    // the calling thread takes part 0
    void run(const std::function<void(int, int)>& f) {
        {
            std::lock_guard<std::mutex> lk(m_);
            job_ = &f;
            pending_ = n_ - 1;
            ++gen_;
        }
        cv_work_.notify_all();
        f(0, n_);
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }
    // WTP_HOST_THREADS, else the host's hardware threads (at most 16: more do not widen faster, the stores saturate the
    // memory system) divided among the ranks that share the host (one process per GPU: `ranks` = the context's world)
    static int default_threads(int ranks = 1) {
        if (const char* e = std::getenv("WTP_HOST_THREADS")) { int v = std::atoi(e); if (v > 0) return v > 64 ? 64 : v; }
        unsigned h = std::thread::hardware_concurrency();
        if (h == 0) h = 4;
        if (ranks > 1) h = h / (unsigned)ranks > 2 ? h / (unsigned)ranks : 2;
        return (int)(h > 16 ? 16 : h);
    }

  private:
    void loop(int id) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int, int)>* f;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_work_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                f = job_;
            }
            (*f)(id, n_);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) cv_done_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(int, int)>* job_ = nullptr;
    int pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

// dst[i] = src[i] zero-extended, i in [0, n): streaming (non-temporal) 16-byte stores once dst is aligned
inline void widen_u32_to_i64(const uint32_t* src, int64_t* dst, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 15u)) { dst[i] = (int64_t)src[i]; ++i; }
    const __m128i zero = _mm_setzero_si128();
    for (; i + 4 <= n; i += 4) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_unpacklo_epi32(v, zero));
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 2), _mm_unpackhi_epi32(v, zero));
    }
    for (; i < n; ++i) dst[i] = (int64_t)src[i];
    _mm_sfence();
}

// The same from 3-byte little-endian values (indices below 2^24 cross PCIe as 3 bytes each): 12 source bytes give 4
// values. Reads 16 bytes per group: the caller keeps 4 readable bytes behind the last group. n is a multiple of 4 except
// possibly in the last call of an array (scalar tail).
__attribute__((target("ssse3"))) inline void widen_u24_to_i64(const unsigned char* src, int64_t* dst, size_t n) {
    size_t i = 0;
    const __m128i zero = _mm_setzero_si128();
    const __m128i pick = _mm_setr_epi8(0, 1, 2, (char)0x80, 3, 4, 5, (char)0x80, 6, 7, 8, (char)0x80, 9, 10, 11, (char)0x80);
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        for (; i + 4 <= n; i += 4) {
            const __m128i v = _mm_shuffle_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 3 * i)), pick);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_unpacklo_epi32(v, zero));
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 2), _mm_unpackhi_epi32(v, zero));
        }
    }
    for (; i < n; ++i) dst[i] = (int64_t)((uint32_t)src[3 * i] | ((uint32_t)src[3 * i + 1] << 8) | ((uint32_t)src[3 * i + 2] << 16));
    _mm_sfence();
}

}  // namespace wtp
