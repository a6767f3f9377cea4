// Input embeddings of the BERT body on packed token rows: out[t,:] = (W[ids[t],:] + T[typ[t],:]) + P[pos[t],:]
// (transformers BertEmbeddings.forward, called from scripts/model/sparse_encoders.py:108), and the scatter of the
// gradient back into the three tables. PyTorch runs three gathers and two adds forward and three sort-based
// embedding_dense_backward pipelines (radix sort + segment kernels, ~25 launches) backward; here it is one gather
// kernel and one scatter kernel. HBM-bound on the fp32 activation ([n, H] read or written once); the scatter uses
// 16-byte vector reductions into the pre-zeroed tables, with the two token-type rows (hit by every token) reduced
// per block in shared memory first.
#include <cuda_runtime.h>

#include <cstdint>

#include "common.h"

namespace sb200 {
namespace {

constexpr int kEmbThreads = 256;
constexpr int kEmbTokens = 32;  // tokens per block

__device__ __forceinline__ void red_add_v4(float* addr, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__global__ void __launch_bounds__(kEmbThreads)
embed_sum_fwd_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ pos, const int64_t* __restrict__ typ,
                     const float* __restrict__ W, const float* __restrict__ P, const float* __restrict__ T, int n, int H,
                     int nW, int nP, int nT, float* __restrict__ out) {
    const int hv = H >> 2;
    const int t0 = blockIdx.x * kEmbTokens;
    const int nt = min(kEmbTokens, n - t0);
    for (int idx = threadIdx.x; idx < nt * hv; idx += kEmbThreads) {
        const int t = t0 + idx / hv, c = (idx % hv) * 4;
        const int64_t iw = min(max(__ldg(ids + t), int64_t(0)), int64_t(nW - 1));
        const int64_t ip = min(max(__ldg(pos + t), int64_t(0)), int64_t(nP - 1));
        const int64_t it = min(max(__ldg(typ + t), int64_t(0)), int64_t(nT - 1));
        const float4 w = __ldg(reinterpret_cast<const float4*>(W + iw * H + c));
        const float4 p = __ldg(reinterpret_cast<const float4*>(P + ip * H + c));
        const float4 y = __ldg(reinterpret_cast<const float4*>(T + it * H + c));
        float4 o;
        o.x = (w.x + y.x) + p.x; o.y = (w.y + y.y) + p.y; o.z = (w.z + y.z) + p.z; o.w = (w.w + y.w) + p.w;
        *reinterpret_cast<float4*>(out + size_t(t) * H + c) = o;
    }
}

__global__ void __launch_bounds__(kEmbThreads)
embed_sum_bwd_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ pos, const int64_t* __restrict__ typ,
                     const float* __restrict__ g, int n, int H, int nW, int nP, int nT, int pad_idx,
                     float* __restrict__ dW, float* __restrict__ dP, float* __restrict__ dT) {
    extern __shared__ float type_acc[];  // [2][H]: rows 0 and 1 of the token-type table (others go straight to HBM)
    const int hv = H >> 2;
    for (int i = threadIdx.x; i < 2 * H; i += kEmbThreads) type_acc[i] = 0.f;
    __syncthreads();
    const int t0 = blockIdx.x * kEmbTokens;
    const int nt = min(kEmbTokens, n - t0);
    for (int idx = threadIdx.x; idx < nt * hv; idx += kEmbThreads) {
        const int t = t0 + idx / hv, c = (idx % hv) * 4;
        const int64_t iw = min(max(__ldg(ids + t), int64_t(0)), int64_t(nW - 1));
        const int64_t ip = min(max(__ldg(pos + t), int64_t(0)), int64_t(nP - 1));
        const int64_t it = min(max(__ldg(typ + t), int64_t(0)), int64_t(nT - 1));
        const float4 v = __ldg(reinterpret_cast<const float4*>(g + size_t(t) * H + c));
        if (iw != pad_idx) red_add_v4(dW + iw * H + c, v);  // torch.nn.Embedding(padding_idx): that row gets no gradient
        red_add_v4(dP + ip * H + c, v);
        if (it < 2) {
            float* a = type_acc + it * H + c;
            atomicAdd(a, v.x); atomicAdd(a + 1, v.y); atomicAdd(a + 2, v.z); atomicAdd(a + 3, v.w);
        } else {
            red_add_v4(dT + it * H + c, v);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < min(nT, 2) * hv; i += kEmbThreads) {
        const float4 v = *reinterpret_cast<const float4*>(type_acc + i * 4);
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) red_add_v4(dT + i * 4, v);
    }
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" int sb200_embed_sum_fwd(const int64_t* ids, const int64_t* pos, const int64_t* typ, const float* W,
                                   const float* P, const float* T, int n, int H, int nW, int nP, int nT, float* out,
                                   sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(ids && pos && typ && W && P && T && out, "embed_sum_fwd: null pointer");
    SB200_REQUIRE(n >= 1 && H >= 4 && H % 4 == 0 && nW >= 1 && nP >= 1 && nT >= 1, "embed_sum_fwd: bad shape n=%d H=%d", n, H);
    embed_sum_fwd_kernel<<<(n + kEmbTokens - 1) / kEmbTokens, kEmbThreads, 0, stream>>>(ids, pos, typ, W, P, T, n, H, nW,
                                                                                        nP, nT, out);
    SB200_CHECK_LAUNCH("embed_sum_fwd_kernel");
    return SB200_OK;
}

extern "C" int sb200_embed_sum_bwd(const int64_t* ids, const int64_t* pos, const int64_t* typ, const float* g, int n,
                                   int H, int nW, int nP, int nT, int pad_idx, float* dW, float* dP,
                                   float* dT, sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(ids && pos && typ && g && dW && dP && dT, "embed_sum_bwd: null pointer");
    SB200_REQUIRE(n >= 1 && H >= 4 && H % 4 == 0 && H <= 4096 && nW >= 1 && nP >= 1 && nT >= 1,
                  "embed_sum_bwd: bad shape n=%d H=%d", n, H);
    embed_sum_bwd_kernel<<<(n + kEmbTokens - 1) / kEmbTokens, kEmbThreads, 2 * H * sizeof(float), stream>>>(
        ids, pos, typ, g, n, H, nW, nP, nT, pad_idx, dW, dP, dT);
    SB200_CHECK_LAUNCH("embed_sum_bwd_kernel");
    return SB200_OK;
}
