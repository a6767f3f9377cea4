// Host-side helpers shared by every translation unit of libsparse_b200.so.
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/sparse_b200.h"

namespace sb200 {

extern thread_local char g_last_error[512];
extern std::atomic<unsigned long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

inline void count_launch(unsigned n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Checks the launch status of the kernel that was just enqueued.
#define SB200_CHECK_LAUNCH(name)                                                                       \
    do {                                                                                               \
        cudaError_t e__ = cudaGetLastError();                                                          \
        if (e__ != cudaSuccess) return ::sb200::fail(SB200_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__)); \
        ::sb200::count_launch();                                                                       \
    } while (0)

#define SB200_CUDA(call)                                                                               \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) return ::sb200::fail(SB200_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

#define SB200_REQUIRE(cond, ...)                                          \
    do {                                                                  \
        if (!(cond)) return ::sb200::fail(SB200_ERR_ARG, __VA_ARGS__);    \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int num_sms();  // SM count of the current device (cached per device)
// Per-device one-shot flags (slot in [0, 64)): returns the previous value and sets the flag.
// Slots: 0/5 head_fwd<1/2>, 1-4 bwd_dw<NCHUNK, bf16>, 16-19 bwd_dw<NCHUNK, fp16>, 6/7 score row kernels, 8-13 colsum<T, NV> instances
// that need > 48 KB of shared memory, 14/15 attn_bwd<64, drop/no drop>, 32/33 attn_fwd<64, drop/no drop>, 20-31 bwd_dw_fit<VB, NCH, bf16|fp16>.
bool device_flag_test_and_set(int slot);

}  // namespace sb200
