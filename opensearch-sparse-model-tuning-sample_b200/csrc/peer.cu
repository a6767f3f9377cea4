// Symmetric peer memory over NVLink for the data-parallel exchange steps of the training step
// (scripts/utils.py:16-23 gather_rep = accelerate.gather; bi_encoder_wrapper.py:130 teacher gather; trainer.py:101-104).
//
// One process per GPU. Every rank allocates the same buffer with cudaMalloc, exports it as a CUDA IPC handle, imports
// the handles of its peers (cudaIpcOpenMemHandle enables peer access over NVLink/NVSwitch) and from then on addresses
// every rank's copy directly from its own kernels:
//   * the fused head epilogue (head_fwd.cu) stores each finished rep[b, v] into the gathered buffer of EVERY rank, so the
//     all-gather of the document vectors rides on the GEMM kernel, tile by tile, instead of following it;
//   * peer_allgather_kernel does the same for small tensors (query token ids, teacher scores, teacher embeddings);
//   * peer_signal_kernel / peer_wait_kernel are the cross-GPU barrier: after its data stores a rank publishes a
//     monotonically increasing epoch in every peer's flag array; consumers spin until all W flags reached their epoch.
// Everything is plain kernels on the caller's stream: CUDA-graph capturable, no host synchronisation, no NCCL.
#include <cuda_runtime.h>

#include "common.h"

namespace sb200 {
namespace {

constexpr int kMaxPeers = 16;

struct PeerPtrs {
    void* p[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(uint32_t* addr, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* addr) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}

// dst[p][rank * n16 + i] = src[i] for every peer p (16-byte vectors); each block walks the source once and fans out.
__global__ void __launch_bounds__(256)
peer_allgather_kernel(const uint4* __restrict__ src, size_t n16, int rank, int world, PeerPtrs dst) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += size_t(gridDim.x) * blockDim.x) {
        const uint4 v = __ldg(src + i);
#pragma unroll 1
        for (int p = 0; p < world; ++p) reinterpret_cast<uint4*>(dst.p[p])[size_t(rank) * n16 + i] = v;
    }
}

// One thread block. epoch = ++(*send_epoch); flags_of_peer[p][rank] = epoch for all p. The kernel boundary in front of
// this launch has completed the data stores of the producing kernel; the system-scope fence + release stores order the
// flag behind them on every link.
__global__ void peer_signal_kernel(uint32_t* __restrict__ send_epoch, int rank, int world, PeerPtrs flags) {
    __shared__ uint32_t e;
    if (threadIdx.x == 0) {
        e = *send_epoch + 1u;
        *send_epoch = e;
    }
    __syncthreads();
    __threadfence_system();
    if (int(threadIdx.x) < world) st_release_sys(static_cast<uint32_t*>(flags.p[threadIdx.x]) + rank, e);
}

// One thread block. epoch = ++(*wait_epoch); spins until every rank's flag in the LOCAL flag array reached it.
__global__ void peer_wait_kernel(uint32_t* __restrict__ wait_epoch, const uint32_t* __restrict__ local_flags, int world) {
    __shared__ uint32_t e;
    if (threadIdx.x == 0) {
        e = *wait_epoch + 1u;
        *wait_epoch = e;
    }
    __syncthreads();
    if (int(threadIdx.x) < world) {
        // epochs wrap after 2^32 gathers; compare with wrap-around arithmetic
        while (int32_t(ld_acquire_sys(local_flags + threadIdx.x) - e) < 0) __nanosleep(40);
    }
    __syncthreads();
    __threadfence_system();
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" int sb200_peer_alloc(size_t bytes, void** ptr) {
    SB200_REQUIRE(ptr != nullptr && bytes > 0, "peer_alloc: bad arguments");
    SB200_CUDA(cudaMalloc(ptr, bytes));
    SB200_CUDA(cudaMemset(*ptr, 0, bytes));
    SB200_CUDA(cudaDeviceSynchronize());
    return SB200_OK;
}

extern "C" int sb200_peer_free(void* ptr) {
    if (ptr != nullptr) SB200_CUDA(cudaFree(ptr));
    return SB200_OK;
}

extern "C" int sb200_peer_export(const void* ptr, void* handle64) {
    SB200_REQUIRE(ptr != nullptr && handle64 != nullptr, "peer_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    SB200_CUDA(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), const_cast<void*>(ptr)));
    return SB200_OK;
}

extern "C" int sb200_peer_import(const void* handle64, void** ptr) {
    SB200_REQUIRE(ptr != nullptr && handle64 != nullptr, "peer_import: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    SB200_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SB200_OK;
}

extern "C" int sb200_peer_close(void* ptr) {
    if (ptr != nullptr) SB200_CUDA(cudaIpcCloseMemHandle(ptr));
    return SB200_OK;
}

static int fill_ptrs(PeerPtrs* out, void* const* ptrs, int world, size_t byte_offset) {
    for (int p = 0; p < kMaxPeers; ++p)
        out->p[p] = p < world ? static_cast<void*>(static_cast<uint8_t*>(ptrs[p]) + byte_offset) : nullptr;
    return SB200_OK;
}

extern "C" int sb200_peer_allgather(const void* src, size_t bytes, int rank, int world, void* const* dst_ptrs,
                                    size_t dst_byte_offset, sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(src && dst_ptrs && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world,
                  "peer_allgather: bad arguments");
    SB200_REQUIRE(bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && dst_byte_offset % 16 == 0,
                  "peer_allgather: 16-byte granularity required (bytes=%zu)", bytes);
    if (bytes == 0) return SB200_OK;
    PeerPtrs dst;
    fill_ptrs(&dst, dst_ptrs, world, dst_byte_offset);
    const size_t n16 = bytes / 16;
    size_t blocks = (n16 + 255) / 256;
    const size_t cap = size_t(4) * num_sms();
    if (blocks > cap) blocks = cap;
    peer_allgather_kernel<<<unsigned(blocks), 256, 0, stream>>>(static_cast<const uint4*>(src), n16, rank, world, dst);
    SB200_CHECK_LAUNCH("peer_allgather_kernel");
    return SB200_OK;
}

extern "C" int sb200_peer_signal(uint32_t* send_epoch, int rank, int world, void* const* flag_ptrs,
                                 size_t flag_byte_offset, sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(send_epoch && flag_ptrs && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world,
                  "peer_signal: bad arguments");
    PeerPtrs flags;
    fill_ptrs(&flags, flag_ptrs, world, flag_byte_offset);
    peer_signal_kernel<<<1, 32, 0, stream>>>(send_epoch, rank, world, flags);
    SB200_CHECK_LAUNCH("peer_signal_kernel");
    return SB200_OK;
}

extern "C" int sb200_peer_wait(uint32_t* wait_epoch, const uint32_t* local_flags, int world, sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(wait_epoch && local_flags && world >= 1 && world <= kMaxPeers, "peer_wait: bad arguments");
    peer_wait_kernel<<<1, 32, 0, stream>>>(wait_epoch, local_flags, world);
    SB200_CHECK_LAUNCH("peer_wait_kernel");
    return SB200_OK;
}
