// HBM-bound vector kernels of the training step: inf-free IDF query lookup (sparse_encoders.py:121-127)
// and the FLOPS / L0-thresholded FLOPS regulariser (trainer.py:61-73), forward and backward.
// All are streaming passes with 128-bit accesses where alignment allows, warp-shuffle reductions, no tensor cores.
#include <cuda_runtime.h>

#include "common.h"

namespace sb200 {
namespace {

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

// ------------------------------------------------------------------------------------------ IDF query
// grid (slabs of kIdfSlab columns, Nq): the block builds its slab of the row in shared memory (zero-fill, then
// relu(idf[id]) for the row's ids that fall into the slab; duplicate ids store the same value, specials and ids outside
// [0, V) are skipped) and streams it out once with 16-byte stores. Every output byte is written exactly once; many
// small blocks keep all SMs busy whatever Nq is (the whole kernel is a Nq*V*4-byte write).
constexpr int kIdfThreads = 256;
constexpr int kIdfSlab = 4096;

template <typename IdT>
__global__ void __launch_bounds__(kIdfThreads)
idf_query_kernel(const IdT* __restrict__ ids, const float* __restrict__ idf, const int32_t* __restrict__ special,
                 int n_special, int Lq, int V, float* __restrict__ q, int32_t* __restrict__ bad_ids) {
    __shared__ __align__(16) float slab[kIdfSlab];
    const int b = blockIdx.y;
    const int s0 = blockIdx.x * kIdfSlab;
    const int s1 = min(V, s0 + kIdfSlab);
    for (int i = threadIdx.x; i < kIdfSlab / 4; i += kIdfThreads)
        reinterpret_cast<float4*>(slab)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    int bad = 0;
    for (int l = threadIdx.x; l < Lq; l += kIdfThreads) {
        const long long id = static_cast<long long>(__ldg(ids + size_t(b) * Lq + l));
        if (id < 0 || id >= V) {
            ++bad;
            continue;
        }
        if (id < s0 || id >= s1) continue;
        bool is_special = false;
        for (int k = 0; k < n_special; ++k) is_special |= (__ldg(special + k) == int32_t(id));
        if (!is_special) slab[int(id) - s0] = fmaxf(__ldg(idf + id), 0.f);
    }
    if (bad_ids != nullptr && blockIdx.x == 0 && bad > 0) atomicAdd(bad_ids, bad);
    __syncthreads();
    // stream the slab out: scalar head up to 16-byte alignment of the destination, float4 body, scalar tail
    float* dst = q + size_t(b) * V + s0;
    const int n = s1 - s0;
    const int head = min(n, int(((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15) >> 2));
    const int body4 = (n - head) >> 2;
    if (int(threadIdx.x) < head) dst[threadIdx.x] = slab[threadIdx.x];
    for (int i = threadIdx.x; i < body4; i += kIdfThreads) {
        const float* sp = slab + head + 4 * i;
        reinterpret_cast<float4*>(dst + head)[i] = make_float4(sp[0], sp[1], sp[2], sp[3]);
    }
    const int tail0 = head + 4 * body4;
    if (tail0 + int(threadIdx.x) < n) dst[tail0 + threadIdx.x] = slab[tail0 + threadIdx.x];
}

// d_idf[v] = sum_b d_q[b,v] * [q[b,v] > 0]
__global__ void __launch_bounds__(256)
idf_query_bwd_kernel(const float* __restrict__ d_q, const float* __restrict__ q, int Nq, int V, float* __restrict__ d_idf) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    float acc = 0.f;
    for (int b = 0; b < Nq; ++b) {
        const size_t o = size_t(b) * V + v;
        if (__ldg(q + o) > 0.f) acc += __ldg(d_q + o);
    }
    d_idf[v] = acc;
}

// ------------------------------------------------------------------------------------------ FLOPS regulariser
// Row pass (only for the L0-thresholded variant or when logging stats are wanted): one block per row -> nnz, rowmask,
// stats. 8-byte loads when the row is 8-byte aligned.
__global__ void __launch_bounds__(256)
flops_rowstat_kernel(const float* __restrict__ rep, int V, float threshold, float* __restrict__ rowmask,
                     float* __restrict__ stats) {
    __shared__ float red[3][8];
    const float* row = rep + size_t(blockIdx.x) * V;
    float nnz = 0.f, psum = 0.f, pmax = 0.f;
    auto take = [&](float x) {
        nnz += (x != 0.f) ? 1.f : 0.f;
        if (x > 0.f) {
            psum += x;
            pmax = fmaxf(pmax, x);
        }
    };
    if ((reinterpret_cast<uintptr_t>(row) & 7) == 0) {
        const int n2 = V >> 1;
        const float2* r2 = reinterpret_cast<const float2*>(row);
#pragma unroll 4
        for (int v = threadIdx.x; v < n2; v += 256) {
            const float2 x = __ldg(r2 + v);
            take(x.x);
            take(x.y);
        }
        if ((V & 1) && threadIdx.x == 0) take(__ldg(row + V - 1));
    } else {
        for (int v = threadIdx.x; v < V; v += 256) take(__ldg(row + v));
    }
    nnz = warp_sum(nnz);
    psum = warp_sum(psum);
    pmax = warp_max(pmax);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        red[0][w] = nnz;
        red[1][w] = psum;
        red[2][w] = pmax;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, s = 0.f, m = 0.f;
        for (int i = 0; i < 8; ++i) {
            a += red[0][i];
            s += red[1][i];
            m = fmaxf(m, red[2][i]);
        }
        // torch.norm(p=0) counts non-zeros; mask = (doc_length > threshold)
        rowmask[blockIdx.x] = (threshold < 0.f || a > threshold) ? 1.f : 0.f;
        if (stats != nullptr) {
            atomicAdd(stats + 0, a);
            atomicAdd(stats + 1, s);
            // entries are >= 0 here, so the int ordering of the bit patterns matches the float ordering
            atomicMax(reinterpret_cast<int*>(stats + 2), __float_as_int(m));
        }
    }
}

// Column pass: colsum[g,v] = sum_n rowmask[n*G+g] * |rep[n,g,v]| and value = sum_c (colsum[c]/N)^2, ONE pass over rep,
// one launch. Each thread owns VEC adjacent columns and walks all N rows with kBatch independent 16-byte loads in
// flight (written out as load-all / use-all: left to itself the compiler serialises the loads on one register set,
// which measured 1.8 TB/s instead of 5.5 TB/s). At any moment the resident blocks read the same few rows across all
// columns, i.e. long contiguous spans of DRAM. The last block to finish adds the per-block shares of the value in block
// order (no float atomics, no memset, deterministic).
constexpr int kColThreads = 128;
__device__ unsigned int g_ticket_flops = 0;

template <int VEC, bool kMasked>
__global__ void __launch_bounds__(kColThreads)
flops_colsum_kernel(const float* __restrict__ rep, const float* __restrict__ rowmask, int N, int GV, int G, int V,
                    float invN, float* __restrict__ colsum, float* __restrict__ partial, float* __restrict__ value) {
    __shared__ float red[kColThreads / 32];
    __shared__ int is_last;
    const int col = (blockIdx.x * kColThreads + threadIdx.x) * VEC;
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    if (col < GV) {
        const int g = col / V;
        auto add = [&](const float* v, float mk) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] += mk * fabsf(v[i]);
        };
        auto load = [&](int n, float* v) {
            const float* p = rep + size_t(n) * GV + col;
            if (VEC == 4) {
                const float4 x = __ldg(reinterpret_cast<const float4*>(p));
                v[0] = x.x; v[1 % VEC] = x.y; v[2 % VEC] = x.z; v[3 % VEC] = x.w;
            } else if (VEC == 2) {
                const float2 x = __ldg(reinterpret_cast<const float2*>(p));
                v[0] = x.x; v[1 % VEC] = x.y;
            } else {
                v[0] = __ldg(p);
            }
        };
        constexpr int kBatch = 8;
        int n = 0;
        for (; n + kBatch <= N; n += kBatch) {
            float x[kBatch][VEC], mk[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                load(n + u, x[u]);
                mk[u] = kMasked ? __ldg(rowmask + size_t(n + u) * G + g) : 1.f;
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) add(x[u], mk[u]);
        }
        for (; n < N; ++n) {
            float x[VEC];
            load(n, x);
            add(x, kMasked ? __ldg(rowmask + size_t(n) * G + g) : 1.f);
        }
    }
    float sq = 0.f;
    if (col < GV) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            colsum[col + i] = acc[i];
            const float mean = acc[i] * invN;
            sq = fmaf(mean, mean, sq);
        }
    }
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < kColThreads / 32; ++i) s += red[i];
        partial[blockIdx.x] = s;
        __threadfence();
        is_last = (atomicAdd(&g_ticket_flops, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    float s = 0.f;
    for (int i = threadIdx.x; i < int(gridDim.x); i += kColThreads) s += __ldcg(partial + i);
    s = warp_sum(s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < kColThreads / 32; ++i) t += red[i];
        *value = t;
        g_ticket_flops = 0u;
    }
}

// d_rep[row, v] (+)= gscale * 2*colsum[g,v]/N^2 * sign(rep) * rowmask[row], rows [row_begin, row_end); rowmask may be
// null (= all ones). 8-byte accesses when everything is 8-byte aligned.
template <int VEC>
__global__ void __launch_bounds__(256)
flops_bwd_kernel(const float* __restrict__ rep, const float* __restrict__ colsum, const float* __restrict__ rowmask,
                 const float* __restrict__ gscale, int G, int V, int row_begin, float k, int accumulate,
                 float* __restrict__ d_rep) {
    const int row = row_begin + blockIdx.y;
    const int g = row % G;
    const float s = __ldg(gscale) * k * (rowmask != nullptr ? __ldg(rowmask + row) : 1.f);
    const float* r = rep + size_t(row) * V;
    const float* cs = colsum + size_t(g) * V;
    float* o = d_rep + size_t(row) * V;
    auto one = [&](float x, float c, float prev) {
        const float sg = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
        const float gval = s * c * sg;
        return accumulate ? (prev + gval) : gval;
    };
    if (VEC == 2) {
        const int n2 = V >> 1;
        for (int v = blockIdx.x * 256 + threadIdx.x; v < n2; v += gridDim.x * 256) {
            const float2 x = __ldg(reinterpret_cast<const float2*>(r) + v);
            const float2 c = __ldg(reinterpret_cast<const float2*>(cs) + v);
            float2 prev = make_float2(0.f, 0.f);
            if (accumulate) prev = reinterpret_cast<const float2*>(o)[v];
            reinterpret_cast<float2*>(o)[v] = make_float2(one(x.x, c.x, prev.x), one(x.y, c.y, prev.y));
        }
    } else {
        for (int v = blockIdx.x * 256 + threadIdx.x; v < V; v += gridDim.x * 256)
            o[v] = one(__ldg(r + v), __ldg(cs + v), accumulate ? o[v] : 0.f);
    }
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" int sb200_idf_query(const void* ids, int ids_elem_bytes, const float* idf, const int32_t* special, int n_special,
                               int Nq, int Lq, int V, float* q, int32_t* bad_ids, sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(ids && idf && q && (special || n_special == 0), "idf_query: null pointer");
    SB200_REQUIRE(ids_elem_bytes == 8 || ids_elem_bytes == 4, "idf_query: ids_elem_bytes=%d", ids_elem_bytes);
    SB200_REQUIRE(Nq >= 1 && Lq >= 1 && V >= 1 && n_special >= 0 && Nq <= 65535, "idf_query: bad shape");
    if (bad_ids != nullptr) SB200_CUDA(cudaMemsetAsync(bad_ids, 0, sizeof(int32_t), stream));
    const dim3 grid((V + kIdfSlab - 1) / kIdfSlab, Nq);
    if (ids_elem_bytes == 8)
        idf_query_kernel<int64_t><<<grid, kIdfThreads, 0, stream>>>(static_cast<const int64_t*>(ids), idf, special,
                                                                    n_special, Lq, V, q, bad_ids);
    else
        idf_query_kernel<int32_t><<<grid, kIdfThreads, 0, stream>>>(static_cast<const int32_t*>(ids), idf, special,
                                                                    n_special, Lq, V, q, bad_ids);
    SB200_CHECK_LAUNCH("idf_query_kernel");
    return SB200_OK;
}

extern "C" int sb200_idf_query_bwd(const float* d_q, const float* q, int Nq, int V, float* d_idf, sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(d_q && q && d_idf && Nq >= 1 && V >= 1, "idf_query_bwd: bad arguments");
    idf_query_bwd_kernel<<<(V + 255) / 256, 256, 0, stream>>>(d_q, q, Nq, V, d_idf);
    SB200_CHECK_LAUNCH("idf_query_bwd_kernel");
    return SB200_OK;
}

extern "C" size_t sb200_flops_workspace_bytes(int N, int G, int V) {
    if (N <= 0 || G <= 0 || V <= 0) return 0;
    return align_up(size_t((size_t(G) * V + kColThreads - 1) / kColThreads) * sizeof(float), 256);   // per-block shares (VEC = 1 worst case)
}

template <int VEC>
static int launch_colsum(const float* rep, const float* mask_in, int N, int GV, int G, int V, float* colsum,
                         float* partial, float* value, cudaStream_t stream) {
    const int blocks = (GV / VEC + kColThreads - 1) / kColThreads;
    if (mask_in != nullptr)
        flops_colsum_kernel<VEC, true><<<blocks, kColThreads, 0, stream>>>(rep, mask_in, N, GV, G, V, 1.f / float(N), colsum,
                                                                          partial, value);
    else
        flops_colsum_kernel<VEC, false><<<blocks, kColThreads, 0, stream>>>(rep, mask_in, N, GV, G, V, 1.f / float(N),
                                                                           colsum, partial, value);
    SB200_CHECK_LAUNCH("flops_colsum_kernel");
    return SB200_OK;
}

extern "C" int sb200_flops_fwd(const float* rep, int N, int G, int V, float threshold, float* colsum, float* rowmask,
                               float* value, float* stats, void* workspace, size_t workspace_bytes,
                               sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(rep && colsum && value, "flops_fwd: null pointer");
    SB200_REQUIRE(N >= 1 && G >= 1 && V >= 1, "flops_fwd: bad shape");
    const long long rows = (long long)N * G;
    const long long GV = (long long)G * V;
    SB200_REQUIRE(rows <= 0x7fffffffLL && GV <= 0x7fffffffLL, "flops_fwd: shape too large");
    if (workspace == nullptr || workspace_bytes < sb200_flops_workspace_bytes(N, G, V))
        return fail(SB200_ERR_WORKSPACE, "flops_fwd: workspace too small");
    const bool row_pass = threshold >= 0.f || stats != nullptr;
    SB200_REQUIRE(!row_pass || rowmask != nullptr, "flops_fwd: rowmask required with a threshold or stats");
    if (row_pass) {
        // row pass: nnz per row -> L0 row mask (+ logging stats). The plain regulariser does not need it.
        if (stats != nullptr) SB200_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(float), stream));
        flops_rowstat_kernel<<<int(rows), 256, 0, stream>>>(rep, V, threshold, rowmask, stats);
        SB200_CHECK_LAUNCH("flops_rowstat_kernel");
    }
    const float* mask_in = (threshold >= 0.f) ? rowmask : nullptr;
    const bool a16 = (reinterpret_cast<uintptr_t>(rep) & 15) == 0;
    float* partial = static_cast<float*>(workspace);
    if (a16 && V % 4 == 0) return launch_colsum<4>(rep, mask_in, N, int(GV), G, V, colsum, partial, value, stream);
    if (a16 && V % 2 == 0) return launch_colsum<2>(rep, mask_in, N, int(GV), G, V, colsum, partial, value, stream);
    return launch_colsum<1>(rep, mask_in, N, int(GV), G, V, colsum, partial, value, stream);
}

extern "C" int sb200_flops_bwd(const float* rep, const float* colsum, const float* rowmask, const float* gscale, int N,
                               int G, int V, int row_begin, int row_end, int accumulate, float* d_rep,
                               sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(rep && colsum && gscale && d_rep, "flops_bwd: null pointer");
    SB200_REQUIRE(N >= 1 && G >= 1 && V >= 1 && row_begin >= 0 && row_end <= N * G && row_begin <= row_end,
                  "flops_bwd: bad shape");
    if (row_end == row_begin) return SB200_OK;
    SB200_REQUIRE(row_end - row_begin <= 65535, "flops_bwd: too many rows");
    const float k = 2.f / (float(N) * float(N));
    const bool vec2 = V % 2 == 0 && ((reinterpret_cast<uintptr_t>(rep) | reinterpret_cast<uintptr_t>(colsum) |
                                      reinterpret_cast<uintptr_t>(d_rep)) & 7) == 0;
    int xb = ((vec2 ? V / 2 : V) + 255) / 256;
    if (xb > 16) xb = 16;
    const dim3 grid(xb, row_end - row_begin);
    if (vec2)
        flops_bwd_kernel<2><<<grid, 256, 0, stream>>>(rep, colsum, rowmask, gscale, G, V, row_begin, k, accumulate, d_rep);
    else
        flops_bwd_kernel<1><<<grid, 256, 0, stream>>>(rep, colsum, rowmask, gscale, G, V, row_begin, k, accumulate, d_rep);
    SB200_CHECK_LAUNCH("flops_bwd_kernel");
    return SB200_OK;
}
