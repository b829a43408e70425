// CPU check of the product's host-side data movement (whatsthepoint.jl_b200/csrc/host_pool.h): the widening of 4-byte and
// 3-byte indices into the caller's int64 table (every alignment, every tail length) and the worker pool. No CUDA needed.
#include <cstdio>
#include <cstring>
#include <atomic>
#include <random>
#include <vector>

#include "../../whatsthepoint.jl_b200/csrc/host_pool.h"

int main() {
    std::mt19937_64 rng(42);
    int bad = 0;
    for (size_t n : {0ul, 1ul, 2ul, 3ul, 4ul, 5ul, 7ul, 8ul, 31ul, 64ul, 1000ul, 4099ul}) {
        for (size_t shift : {0ul, 1ul}) {                       // destination 16-byte aligned or not
            std::vector<uint32_t> src(n + 8);
            for (auto& v : src) v = (uint32_t)rng();
            std::vector<int64_t> dst(n + 4, -7), want(n + 4, -7);
            for (size_t i = 0; i < n; ++i) want[shift + i] = (int64_t)src[i];
            wtp::widen_u32_to_i64(src.data(), dst.data() + shift, n);
            if (dst != want) { std::printf("widen_u32 n=%zu shift=%zu differs\n", n, shift); ++bad; }
            // 3-byte packing as pack24_kernel writes it: value i at bytes 3i .. 3i+2, little endian
            std::vector<unsigned char> packed(3 * n + 16, 0xee);
            for (size_t i = 0; i < n; ++i) { const uint32_t v = src[i] & 0xffffffu; packed[3 * i] = v & 255; packed[3 * i + 1] = (v >> 8) & 255; packed[3 * i + 2] = v >> 16; }
            std::vector<int64_t> dst3(n + 4, -7), want3(n + 4, -7);
            for (size_t i = 0; i < n; ++i) want3[shift + i] = (int64_t)(src[i] & 0xffffffu);
            wtp::widen_u24_to_i64(packed.data(), dst3.data() + shift, n);
            if (dst3 != want3) { std::printf("widen_u24 n=%zu shift=%zu differs\n", n, shift); ++bad; }
        }
    }
    for (int threads : {1, 2, 5}) {
        wtp::HostPool pool(threads);
        for (int rep = 0; rep < 50; ++rep) {
            std::atomic<int> sum{0}, calls{0};
            pool.run([&](int part, int parts) { sum += part; ++calls; if (parts != threads) sum += 1000; });
            if (calls != threads || sum != threads * (threads - 1) / 2) { std::printf("pool threads=%d rep=%d: calls=%d sum=%d\n", threads, rep, calls.load(), sum.load()); ++bad; }
        }
    }
    if (wtp::HostPool::default_threads(1) < 1 || wtp::HostPool::default_threads(8) < 2 || wtp::HostPool::default_threads(8) > wtp::HostPool::default_threads(1)) { std::printf("default_threads\n"); ++bad; }
    std::printf(bad ? "FAILED\n" : "OK\n");
    return bad ? 1 : 0;
}
