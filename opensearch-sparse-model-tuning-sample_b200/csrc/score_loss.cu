// In-batch query x doc scoring and the ranking losses of scripts/train/loss.py, plus the encode-output
// compaction (sparse_encoders.py:137-150, 178-179) and the teacher min-max normalisation
// (bi_encoder_wrapper.py:133-138). fp32 CUDA-core arithmetic throughout (the reference runs these GEMMs in
// fp32 with TF32 off); the score kernels are HBM/L2-bound: (Nq+Nd)*V*4 bytes against Nq*Nd*V*2 flop.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "common.h"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace sb200 {
namespace {

__host__ __device__ inline size_t align_up_sz(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------ block reductions
template <int THREADS>
__device__ __forceinline__ float block_sum(float x, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) s += red[i];
    return s;
}
template <int THREADS>
__device__ __forceinline__ float block_max(float x, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    float s = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) s = fmaxf(s, red[i]);
    return s;
}
template <int THREADS>
__device__ __forceinline__ float block_min(float x, float* red) {
    return -block_max<THREADS>(-x, red);
}

// Deterministic "last block finishes the sum": every block publishes its partial results, takes a ticket, and the
// block that draws the last ticket adds the n partials in a fixed order and resets the ticket for the next launch.
// The ticket counters are per-device globals: one in-flight launch per kernel and device (stream-ordered use).
template <int THREADS>
__device__ __forceinline__ void finish_sum_last_block(unsigned int* ticket, unsigned int n_blocks, const float* partial,
                                                      int n, float scale, float* out, float* red) {
    __shared__ int is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == n_blocks - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += THREADS) acc += __ldcg(partial + i);
    acc = block_sum<THREADS>(acc, red);
    if (threadIdx.x == 0) {
        *out = acc * scale;
        *ticket = 0u;
    }
}

__device__ unsigned int g_ticket_rank_loss = 0;
__device__ unsigned int g_ticket_fused = 0;
__device__ unsigned int g_ticket_group = 0;

// ------------------------------------------------------------------------------------------ scores, in-batch
// Dense fallback. S[i,j] += sum_{k in split} q[i,k] d[j,k].  64x64 output tile per block, 4x4 per thread, K staged 32
// at a time through shared memory (transposed, stride 65 -> conflict-free), split-K over blockIdx.z merged with atomics.
constexpr int kST = 64, kSK = 32, kSStride = kST + 1;

__global__ void __launch_bounds__(256)
scores_tile_kernel(const float* __restrict__ q, const float* __restrict__ d, int Nq, int Nd, int V, int kchunk,
                   const int* __restrict__ dense_flag, float* __restrict__ S) {
    __shared__ float qs[kSK * kSStride];
    __shared__ float ds[kSK * kSStride];
    if (dense_flag != nullptr && *dense_flag == 0) return;  // the sparse-query kernel produced S
    const int i0 = blockIdx.y * kST, j0 = blockIdx.x * kST;
    const int k_begin = blockIdx.z * kchunk;
    const int k_end = min(V, k_begin + kchunk);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int lk = threadIdx.x & 31, lr = threadIdx.x >> 5;  // loader: k offset, row offset
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    for (int k0 = k_begin; k0 < k_end; k0 += kSK) {
        const int k = k0 + lk;
        const bool k_ok = k < k_end;
        float qv[8], dv[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int r = lr + 8 * p;
            qv[p] = (k_ok && i0 + r < Nq) ? __ldg(q + size_t(i0 + r) * V + k) : 0.f;
            dv[p] = (k_ok && j0 + r < Nd) ? __ldg(d + size_t(j0 + r) * V + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int r = lr + 8 * p;
            qs[lk * kSStride + r] = qv[p];
            ds[lk * kSStride + r] = dv[p];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kSK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                a[e] = qs[kk * kSStride + ty * 4 + e];
                b[e] = ds[kk * kSStride + tx + 16 * e];
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
        }
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        const int i = i0 + ty * 4 + x;
        if (i >= Nq) continue;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int j = j0 + tx + 16 * y;
            if (j < Nd) {
                if (gridDim.z == 1) S[size_t(i) * Nd + j] = acc[x][y];
                else atomicAdd(S + size_t(i) * Nd + j, acc[x][y]);
            }
        }
    }
}

// Dense fallback on the legacy tensor-core path: the same 64x64 tile / split-K scheme, but the inner product runs as
// mma.sync.m16n8k8 TF32 with the 3xTF32 split (x = hi + lo, both TF32; lo*hi + hi*lo + hi*hi, small terms first), which
// keeps fp32-level accuracy (the reference multiplies in fp32 with TF32 off). 8 warps = 4 (query rows) x 2 (doc rows);
// a warp owns a 16 x 32 patch = 4 MMA n-tiles. Shared-memory rows are padded to 36 floats: every fragment load of a warp
// hits 32 different banks. Needs 8-byte aligned rows (V even); other shapes use the CUDA-core kernel above.
constexpr int kMmaStride = kSK + 4;

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float rest = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rest));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(256)
scores_tile_mma_kernel(const float* __restrict__ q, const float* __restrict__ d, int Nq, int Nd, int V, int kchunk,
                       const int* __restrict__ dense_flag, float* __restrict__ S) {
    __shared__ __align__(16) float qs[kST * kMmaStride];
    __shared__ __align__(16) float ds[kST * kMmaStride];
    if (dense_flag != nullptr && *dense_flag == 0) return;  // the sparse-query kernel produced S
    const int i0 = blockIdx.y * kST, j0 = blockIdx.x * kST;
    const int k_begin = blockIdx.z * kchunk;
    const int k_end = min(V, k_begin + kchunk);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wr = warp >> 1, wc = warp & 1;        // 4 x 2 warps
    const int g = lane >> 2, t = lane & 3;
    float acc[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;

    for (int k0 = k_begin; k0 < k_end; k0 += kSK) {
        float2 qv[4], dv[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {               // 64 rows x 16 float2 per operand, 4 per thread
            const int idx = threadIdx.x + 256 * p;
            const int r = idx >> 4, k = k0 + 2 * (idx & 15);
            const bool k_ok = k < k_end;            // k_end is even (V even, kchunk a multiple of 32)
            qv[p] = (k_ok && i0 + r < Nq) ? __ldg(reinterpret_cast<const float2*>(q + size_t(i0 + r) * V + k))
                                          : make_float2(0.f, 0.f);
            dv[p] = (k_ok && j0 + r < Nd) ? __ldg(reinterpret_cast<const float2*>(d + size_t(j0 + r) * V + k))
                                          : make_float2(0.f, 0.f);
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int idx = threadIdx.x + 256 * p;
            const int r = idx >> 4, c = 2 * (idx & 15);
            *reinterpret_cast<float2*>(&qs[r * kMmaStride + c]) = qv[p];
            *reinterpret_cast<float2*>(&ds[r * kMmaStride + c]) = dv[p];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kSK / 8; ++kk) {
            uint32_t a_hi[4], a_lo[4];
            const float* ap = qs + (wr * 16 + g) * kMmaStride + kk * 8 + t;
            split_tf32(ap[0], a_hi[0], a_lo[0]);
            split_tf32(ap[8 * kMmaStride], a_hi[1], a_lo[1]);
            split_tf32(ap[4], a_hi[2], a_lo[2]);
            split_tf32(ap[8 * kMmaStride + 4], a_hi[3], a_lo[3]);
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                uint32_t b_hi[2], b_lo[2];
                const float* bp = ds + (wc * 32 + n * 8 + g) * kMmaStride + kk * 8 + t;
                split_tf32(bp[0], b_hi[0], b_lo[0]);
                split_tf32(bp[4], b_hi[1], b_lo[1]);
                mma_tf32(acc[n], a_lo, b_hi);
                mma_tf32(acc[n], a_hi, b_lo);
                mma_tf32(acc[n], a_hi, b_hi);
            }
        }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = i0 + wr * 16 + g + ((e & 2) ? 8 : 0);
            const int j = j0 + wc * 32 + n * 8 + 2 * t + (e & 1);
            if (i < Nq && j < Nd) {
                if (gridDim.z == 1) S[size_t(i) * Nd + j] = acc[n][e];
                else atomicAdd(S + size_t(i) * Nd + j, acc[n][e]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ ranking-loss row
// Loss contribution of query row i (already scaled by the batch mean) and, if g != nullptr, d loss / d S[i, :].
// Result valid in thread 0. All THREADS threads of the block must call it. loss.py:33-42, 64-76, 94-106.
template <int THREADS>
__device__ float rank_loss_row(int mode, const float* __restrict__ s, const float* __restrict__ t, float* __restrict__ g,
                               int i, int Nq, int C, int G, int in_batch, float invT, float* red) {
    const float invNq = 1.f / float(Nq);
    if (mode == SB200_LOSS_INFONCE) {
        // selected columns: own positive + every hard negative (loss.py:90-101); own docs only when !in_batch
        const int pos = in_batch ? i * G : 0;
        auto selected = [&](int j) { return !in_batch || j == pos || (j % G) != 0; };
        float m = -CUDART_INF_F;
        for (int j = threadIdx.x; j < C; j += THREADS)
            if (selected(j)) m = fmaxf(m, s[j]);
        m = block_max<THREADS>(m, red);
        float z = 0.f;
        for (int j = threadIdx.x; j < C; j += THREADS)
            if (selected(j)) z += expf(s[j] - m);
        z = block_sum<THREADS>(z, red);
        const float lse = m + logf(z);
        if (g != nullptr) {
            for (int j = threadIdx.x; j < C; j += THREADS) {
                float v = 0.f;
                if (selected(j)) v = (expf(s[j] - lse) - (j == pos ? 1.f : 0.f)) * invNq;
                g[j] = v;
            }
        }
        return (lse - s[pos]) * invNq;
    }
    if (mode == SB200_LOSS_KLDIV) {
        float ms = -CUDART_INF_F, mt = -CUDART_INF_F;
        for (int j = threadIdx.x; j < C; j += THREADS) {
            ms = fmaxf(ms, s[j] * invT);
            mt = fmaxf(mt, t[j] * invT);
        }
        ms = block_max<THREADS>(ms, red);
        mt = block_max<THREADS>(mt, red);
        float zs = 0.f, zt = 0.f;
        for (int j = threadIdx.x; j < C; j += THREADS) {
            zs += expf(s[j] * invT - ms);
            zt += expf(t[j] * invT - mt);
        }
        zs = block_sum<THREADS>(zs, red);
        zt = block_sum<THREADS>(zt, red);
        const float lzs = logf(zs), lzt = logf(zt);
        float acc = 0.f;
        for (int j = threadIdx.x; j < C; j += THREADS) {
            const float lps = s[j] * invT - ms - lzs;
            const float lpt = t[j] * invT - mt - lzt;
            const float pt = expf(lpt);
            if (pt > 0.f) acc += pt * (lpt - lps);  // xlogy(t,t) - t*input
            if (g != nullptr) g[j] = (expf(lps) - pt) * invT * invNq;
        }
        acc = block_sum<THREADS>(acc, red);
        return acc * invNq;
    }
    // margin MSE: margins against column 0 (loss.py:52-55)
    const float s0 = s[0] * invT, t0 = t[0] * invT;
    const float norm = 1.f / (float(Nq) * float(C - 1));
    float acc = 0.f, g0 = 0.f;
    for (int j = 1 + threadIdx.x; j < C; j += THREADS) {
        const float diff = (s0 - s[j] * invT) - (t0 - t[j] * invT);
        acc = fmaf(diff, diff, acc);
        const float gj = 2.f * diff * norm * invT;
        g0 += gj;
        if (g != nullptr) g[j] = -gj;
    }
    acc = block_sum<THREADS>(acc, red);
    g0 = block_sum<THREADS>(g0, red);
    if (threadIdx.x == 0 && g != nullptr) g[0] = g0;
    return acc * norm;
}

// One block per query row; the last block adds the row losses in row order (deterministic, no memset).
__global__ void __launch_bounds__(256)
rank_loss_kernel(int mode, const float* __restrict__ S, const float* __restrict__ teacher, int Nq, int C, int G,
                 int in_batch, float invT, float* __restrict__ loss, float* __restrict__ dS, float* __restrict__ rowloss) {
    __shared__ float red[8];
    const int i = blockIdx.x;
    const float v = rank_loss_row<256>(mode, S + size_t(i) * C, teacher ? teacher + size_t(i) * C : nullptr,
                                       dS ? dS + size_t(i) * C : nullptr, i, Nq, C, G, in_batch, invT, red);
    if (threadIdx.x == 0) rowloss[i] = v;
    finish_sum_last_block<256>(&g_ticket_rank_loss, gridDim.x, rowloss, Nq, 1.f, loss, red);
}

// ------------------------------------------------------------------------------------------ scores, own docs only
// S[i,g] = q[i,:] . d[i*G+g,:] and, fused, the ranking loss of row i. A thread-block CLUSTER of `KS` CTAs shares one
// query: each CTA streams its slice of the vocabulary (q once, the G doc rows once), the partial dot products meet in
// the leader's shared memory through DSMEM, the leader writes S[i,:], evaluates the loss row and takes part in the
// deterministic final sum. One launch, no atomics on S, no memset.
constexpr int kGroupThreads = 256;
constexpr int kGroupMaxG = 64;

__global__ void __launch_bounds__(kGroupThreads)
score_group_kernel(const float* __restrict__ q, const float* __restrict__ d, int Nq, int G, int V, int mode,
                   const float* __restrict__ teacher, float invT, float* __restrict__ S, float* __restrict__ dS,
                   float* __restrict__ rowloss, float* __restrict__ loss) {
    __shared__ float red[kGroupThreads / 32];
    __shared__ float part[kGroupMaxG];        // this CTA's partial dot products
    __shared__ float srow[kGroupMaxG];        // leader: the finished score row
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int ks = cluster.num_blocks(), kr = cluster.block_rank();
    const int i = blockIdx.y;
    const int kchunk = int(align_up_sz(size_t((V + ks - 1) / ks), 4));
    const int k_begin = min(V, int(kr) * kchunk), k_end = min(V, k_begin + kchunk);
    const float* qi = q + size_t(i) * V;
    for (int g0 = 0; g0 < G; g0 += 8) {
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        for (int k = k_begin + threadIdx.x; k < k_end; k += kGroupThreads) {
            const float qv = __ldg(qi + k);
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (g0 + e < G) acc[e] = fmaf(qv, __ldg(d + (size_t(i) * G + g0 + e) * V + k), acc[e]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (g0 + e < G) {  // block-uniform
                const float s = block_sum<kGroupThreads>(acc[e], red);
                if (threadIdx.x == 0) part[g0 + e] = s;
            }
        }
    }
    cluster.sync();
    if (kr == 0) {
        for (int g = threadIdx.x; g < G; g += kGroupThreads) {
            float s = part[g];
            for (unsigned int r = 1; r < ks; ++r) s += *cluster.map_shared_rank(&part[g], r);   // fixed order
            srow[g] = s;
            S[size_t(i) * G + g] = s;
        }
    }
    cluster.sync();   // peers' shared memory must stay alive until the leader has read it
    if (kr != 0 || mode < 0) return;
    __syncthreads();
    const float v = rank_loss_row<kGroupThreads>(mode, srow, teacher ? teacher + size_t(i) * G : nullptr,
                                                 dS ? dS + size_t(i) * G : nullptr, i, Nq, G, G, 0, invT, red);
    if (threadIdx.x == 0) rowloss[i] = v;
    finish_sum_last_block<kGroupThreads>(&g_ticket_group, unsigned(Nq), rowloss, Nq, 1.f, loss, red);
}

// ------------------------------------------------------------------------------------------ sparse-query path
// Queries are very sparse (inf-free: at most Lq token ids; learned: tens to hundreds of entries once trained), so
// q.d^T = sum over the query's non-zeros of val * d[j, col]. The query rows are thresholded (!= 0) into ordered
// (col, val) lists once (q_compact_kernel); then every DOCUMENT row is streamed through shared memory exactly once and
// all query lists are gathered from it: HBM traffic = one pass over d, S[i,j] written once, no atomics (deterministic).
//
// The row kernel is persistent (one CTA per SM, rows round-robin) and keeps the memory pipe busy with a ring of
// shared-memory stages, each holding one PART of a row (a row is cut into n_parts equal column ranges so that any
// vocabulary size fits): a part is fetched with one bulk async copy (TMA engine, mbarrier completion), two parts in
// flight per SM while the current one is consumed.
// Two gather modes, chosen on the device from the list lengths (block-uniform):
//   (A) registers: tpq = 512/Nq threads share a query, each keeps <= 32 entries in registers for the whole kernel;
//   (B) streamed lists: warps walk the queries, lanes the (ascending) entries of the current part, from L2.
// A query with more than kQCap non-zeros raises the device-side flag: the dense kernels then do the work.
constexpr int kQCap = 1024;
constexpr int kRowThreads = 512;
constexpr int kEPT = 32;
constexpr int kRowStages = 3;
constexpr int kRowSmemBudget = 200 * 1024;   // dynamic shared memory for the stages
constexpr int kRowMaxQ = 2048;               // mode (B): cursor + partial score per query in shared memory

struct QLists {
    int* flag;      // [1]  written by the forward row kernel: 1 = some row has more than kQCap entries -> dense path
    int* nnz;       // [Nq] true non-zero count per query row
    int* cols;      // [Nq][kQCap] ascending
    float* vals;    // [Nq][kQCap]
};

__host__ __device__ inline size_t qlists_bytes(int Nq) {
    return 256 + align_up_sz(size_t(Nq) * 4, 256) + 2 * align_up_sz(size_t(Nq) * kQCap * 4, 256);
}
inline QLists qlists_carve(void* ws, int Nq) {
    uint8_t* p = static_cast<uint8_t*>(ws);
    QLists q;
    q.flag = reinterpret_cast<int*>(p);
    q.nnz = reinterpret_cast<int*>(p + 256);
    q.cols = reinterpret_cast<int*>(p + 256 + align_up_sz(size_t(Nq) * 4, 256));
    q.vals = reinterpret_cast<float*>(p + 256 + align_up_sz(size_t(Nq) * 4, 256) + align_up_sz(size_t(Nq) * kQCap * 4, 256));
    return q;
}

// Ordered stream compaction of one row by one block (ballots + block scan), no atomics, no pre-zeroed counters. A
// thread owns 2 adjacent columns per load (rows are 8-byte aligned when V is even) and keeps kCompactBatch loads in
// flight, so a 30522-column row is one (1024 threads) or two (512 threads) round trips to memory.
// store(offset, column, value) is called for every non-zero in ascending column order; returns the row's count.
constexpr int kCompactThreads = 1024;
constexpr int kCompactBatch = 16;  // independent 8-byte loads in flight per thread

template <int THREADS, typename Store>
__device__ __forceinline__ int compact_row(const float* __restrict__ row, int V, int (*warp_cnt)[32], Store store) {
    constexpr int kWarps = THREADS / 32;
    const bool vec2 = (reinterpret_cast<uintptr_t>(row) & 7) == 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int base = 0;
    for (int v0 = 0; v0 < V; v0 += THREADS * kCompactBatch * 2) {
        float2 x[kCompactBatch];
        uint32_t b0[kCompactBatch], b1[kCompactBatch];
#pragma unroll
        for (int b = 0; b < kCompactBatch; ++b) {
            const int v = v0 + (b * THREADS + threadIdx.x) * 2;
            if (vec2 && v + 1 < V) x[b] = __ldg(reinterpret_cast<const float2*>(row + v));
            else x[b] = make_float2(v < V ? __ldg(row + v) : 0.f, v + 1 < V ? __ldg(row + v + 1) : 0.f);
        }
#pragma unroll
        for (int b = 0; b < kCompactBatch; ++b) {
            b0[b] = __ballot_sync(0xffffffffu, x[b].x != 0.f);
            b1[b] = __ballot_sync(0xffffffffu, x[b].y != 0.f);
            if (lane == 0) warp_cnt[b][warp] = __popc(b0[b]) + __popc(b1[b]);
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < kCompactBatch; ++b) {
            const int c = (lane < kWarps) ? warp_cnt[b][lane] : 0;     // lane w holds the count of warp w
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const int before = __shfl_sync(0xffffffffu, incl - c, warp);   // entries of lower warps
            const uint32_t lt = (1u << lane) - 1u;
            int off = base + before + __popc(b0[b] & lt) + __popc(b1[b] & lt);   // both columns of lower lanes first
            const int v = v0 + (b * THREADS + threadIdx.x) * 2;
            if (x[b].x != 0.f) store(off++, v, x[b].x);
            if (x[b].y != 0.f) store(off, v + 1, x[b].y);
            base += total;
        }
        __syncthreads();
    }
    return base;
}

__global__ void __launch_bounds__(kCompactThreads)
q_compact_kernel(const float* __restrict__ q, int V, QLists L) {
    __shared__ int warp_cnt[kCompactBatch][32];
    const int i = blockIdx.x;
    const int n = compact_row<kCompactThreads>(q + size_t(i) * V, V, warp_cnt, [&](int off, int v, float x) {
        if (off < kQCap) {
            L.cols[size_t(i) * kQCap + off] = v;
            L.vals[size_t(i) * kQCap + off] = x;
        }
    });
    if (threadIdx.x == 0) L.nnz[i] = n;
}

struct LossArgs {
    int mode;                 // SB200_LOSS_* or -1: scores only
    const float* teacher;     // [Nq, Nd] or nullptr
    int G;
    float invT;
    float* loss;              // [1]
    float* dS;                // [Nq, Nd] or nullptr
    float* rowloss;           // [Nq] scratch
};

// Fetches part `p` of document row `j` into `stage` with ONE bulk async copy (TMA engine, mbarrier completion): called
// by warp 0; lane 0 drives the copy of the 16-byte aligned interior, lanes 1-6 move the <= 3 floats in front of / behind
// it with ordinary loads. The stage keeps the source's phase inside a 16-byte line (element x of the part lives at
// stage[a + x]), so source and destination are co-aligned. No LSU issue slots are spent on the transfer: an A/B run with
// 16-byte cp.async issued by all threads (same ring) measured 78.6 us against 59.3 us for this version at the large
// shape -- the copies' issue and shared-memory write slots compete with the gather's shared-memory loads.
__device__ __forceinline__ void row_part_issue_bulk(const float* __restrict__ d, int V, int j, int p, int part_len,
                                                    float* stage, uint64_t* bar, int lane) {
    const int c0 = p * part_len;
    const int len = min(V, c0 + part_len) - c0;
    const float* src = d + size_t(j) * V + c0;
    const int a = int(reinterpret_cast<uintptr_t>(src) & 15) >> 2;
    const int head = min(len, (4 - a) & 3);
    const int nbulk = ((len - head) >> 2) << 2;
    const int tail0 = head + nbulk;
    if (lane == 0) {
        if (nbulk > 0) {
            mbar_arrive_expect_tx(bar, uint32_t(nbulk) * 4u);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(stage + head + a)), "l"(src + head), "r"(uint32_t(nbulk) * 4u), "r"(smem_u32(bar))
                         : "memory");
        } else {
            mbar_arrive(bar);
        }
    } else if (lane <= 3) {
        if (lane - 1 < head) stage[a + lane - 1] = __ldg(src + lane - 1);
    } else if (lane <= 6) {
        const int t = tail0 + lane - 4;
        if (t < len) stage[a + t] = __ldg(src + t);
    }
}
template <bool kFused>
__global__ void __launch_bounds__(kRowThreads, 1)
scores_docrow_kernel(const float* __restrict__ d, int Nq, int Nd, int V, int n_parts, int part_len, QLists L,
                     float* __restrict__ S, LossArgs la) {
    extern __shared__ __align__(128) float stages[];            // [kRowStages][part_len + 8]
    __shared__ __align__(8) uint64_t full_bar[kRowStages];      // completion barriers of the stages
    __shared__ int cursor_s[kRowMaxQ];                          // mode (B)
    __shared__ float acc_s[kRowMaxQ];
    __shared__ float red[kRowThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stage_floats = part_len + 8;

    // ---- list lengths -> overflow flag and gather mode (block-uniform)
    int mx = 0;
    for (int i = threadIdx.x; i < Nq; i += kRowThreads) mx = max(mx, __ldg(L.nnz + i));
    mx = int(block_max<kRowThreads>(float(mx), red));
    const bool overflow = mx > kQCap;
    if (blockIdx.x == 0 && threadIdx.x == 0) *L.flag = overflow ? 1 : 0;
    int tpq = 32;                                               // threads sharing one query in mode (A)
    while (tpq > 1 && kRowThreads / tpq < Nq) tpq >>= 1;
    const bool mode_a = (kRowThreads / tpq >= Nq) && (mx <= kEPT * tpq);
    const bool mode_b_ok = Nq <= kRowMaxQ;
    const bool work = !overflow && (mode_a || mode_b_ok);       // otherwise the dense kernels produce S
    if (blockIdx.x == 0 && threadIdx.x == 0 && !work) *L.flag = 1;

    if (work) {
        int col[kEPT];
        float val[kEPT];
        const int qi = threadIdx.x / tpq, sub = threadIdx.x % tpq;
        if (mode_a) {
            const int n = (qi < Nq) ? __ldg(L.nnz + qi) : 0;
#pragma unroll
            for (int e = 0; e < kEPT; ++e) {
                const int k = sub + e * tpq;
                const bool ok = k < n;
                col[e] = ok ? __ldg(L.cols + size_t(qi) * kQCap + k) : 0x7fffffff;
                val[e] = ok ? __ldg(L.vals + size_t(qi) * kQCap + k) : 0.f;
            }
        }
        const int my_rows = (Nd - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
        const int total = my_rows * n_parts;
        if (threadIdx.x == 0) {
            for (int s = 0; s < kRowStages; ++s) mbar_init(&full_bar[s], 1);
            fence_mbar_init();
        }
        __syncthreads();
        auto issue = [&](int t) {
            if (warp == 0)
                row_part_issue_bulk(d, V, blockIdx.x + (t / n_parts) * gridDim.x, t % n_parts, part_len,
                                    stages + (t % kRowStages) * stage_floats, &full_bar[t % kRowStages], lane);
        };
        for (int t = 0; t < kRowStages - 1 && t < total; ++t) issue(t);   // prologue: kRowStages-1 parts in flight
        float acc = 0.f;
        for (int t = 0; t < total; ++t) {
            const int s = t % kRowStages;
            const int r = t / n_parts, p = t - r * n_parts;
            const int j = blockIdx.x + r * gridDim.x;
            const int c0 = p * part_len;
            const int plen = min(V, c0 + part_len) - c0;
            const float* st = stages + s * stage_floats +
                              (int(reinterpret_cast<uintptr_t>(d + size_t(j) * V + c0) & 15) >> 2);
            mbar_wait(&full_bar[s], uint32_t(t / kRowStages) & 1u);   // the bulk copy of part t has landed
            __syncthreads();     // its scalar head / tail floats are visible; everybody is done gathering part t-1
            if (t + kRowStages - 1 < total) issue(t + kRowStages - 1);   // refill the stage part t-1 lived in
            if (mode_a) {
#pragma unroll
                for (int e = 0; e < kEPT; ++e) {
                    const unsigned int c = unsigned(col[e] - c0);
                    if (c < unsigned(plen)) acc = fmaf(val[e], st[c], acc);
                }
                if (p == n_parts - 1) {
                    for (int o = tpq >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (sub == 0 && qi < Nq) S[size_t(qi) * Nd + j] = acc;
                    acc = 0.f;
                }
            } else {
                const int c1 = c0 + plen;
                for (int i = warp; i < Nq; i += kRowThreads / 32) {
                    const int n = min(__ldg(L.nnz + i), kQCap);
                    int k0 = (p == 0) ? 0 : cursor_s[i];
                    float a = 0.f;
                    while (true) {
                        const int k = k0 + lane;
                        const int c = (k < n) ? __ldg(L.cols + size_t(i) * kQCap + k) : 0x7fffffff;
                        const bool in = c < c1;
                        if (in) a = fmaf(__ldg(L.vals + size_t(i) * kQCap + k), st[c - c0], a);
                        const int cnt = __popc(__ballot_sync(0xffffffffu, in));
                        k0 += cnt;
                        if (cnt < 32) break;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                    if (lane == 0) {
                        if (p > 0) a += acc_s[i];
                        if (p == n_parts - 1) S[size_t(i) * Nd + j] = a;
                        else { acc_s[i] = a; cursor_s[i] = k0; }
                    }
                }
            }
        }
    }
    if constexpr (kFused) {
        // ---- ranking loss on the finished score matrix: grid-wide barrier (cooperative launch), rows over blocks
        __threadfence();
        cg::this_grid().sync();
        if (!work) return;     // grid-uniform: the separate dense kernels + rank_loss_kernel follow
        for (int i = blockIdx.x; i < Nq; i += gridDim.x) {
            const float v = rank_loss_row<kRowThreads>(la.mode, S + size_t(i) * Nd,
                                                       la.teacher ? la.teacher + size_t(i) * Nd : nullptr,
                                                       la.dS ? la.dS + size_t(i) * Nd : nullptr, i, Nq, Nd, la.G, 1,
                                                       la.invT, red);
            if (threadIdx.x == 0) la.rowloss[i] = v;
            __syncthreads();
        }
        finish_sum_last_block<kRowThreads>(&g_ticket_fused, gridDim.x, la.rowloss, Nq, 1.f, la.loss, red);
    }
}

// ------------------------------------------------------------------------------------------ very sparse queries
// Direct gather: one block per query row i, ONE launch for scores + loss + dS.
//   (1) ordered compaction of q[i, :] into shared memory (and into the global lists the backward kernel reuses);
//   (2) warps walk the documents j (four at a time), lanes the entries: S[i,j] = sum_k val_k * d[j, col_k], read
//       straight from global memory -- Nq * nnz 32-byte sectors per document instead of the V * 4 bytes of streaming the
//       row, the better deal while Nq * nnz * 8 <= V (inf-free queries of a per-GPU batch);
//   (3) the ranking-loss row of query i from the score row held in shared memory;
//   (4) the deterministic last-block sum of the row losses.
constexpr int kGatherThreads = 512;
constexpr int kGatherCap = 256;       // entries per query row kept in shared memory
constexpr int kGatherMaxNd = 8192;    // score row kept in shared memory
__device__ unsigned int g_ticket_gather = 0;

__global__ void __launch_bounds__(kGatherThreads)
score_gather_kernel(const float* __restrict__ q, const float* __restrict__ d, int Nq, int Nd, int V, QLists L,
                    float* __restrict__ S, LossArgs la) {
    __shared__ int col_s[kGatherCap];
    __shared__ float val_s[kGatherCap];
    __shared__ int warp_cnt[kCompactBatch][32];
    __shared__ float red[kGatherThreads / 32];
    extern __shared__ float srow[];     // [Nd]
    const int i = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kGatherThreads / 32;
    // ---- (1) ordered compaction
    const int base = compact_row<kGatherThreads>(q + size_t(i) * V, V, warp_cnt, [&](int off, int v, float x) {
        if (off < kGatherCap) {
            col_s[off] = v;
            val_s[off] = x;
        }
        if (off < kQCap) {
            L.cols[size_t(i) * kQCap + off] = v;
            L.vals[size_t(i) * kQCap + off] = x;
        }
    });
    if (threadIdx.x == 0) {
        L.nnz[i] = base;
        if (i == 0) *L.flag = 0;
    }
    __syncthreads();
    const int n = min(base, kGatherCap);   // the host only picks this kernel when the caller's bound fits
    // ---- (2) gather
    for (int j0 = warp * 4; j0 < Nd; j0 += kWarps * 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = lane; k < n; k += 32) {
            const int c = col_s[k];
            const float w = val_s[k];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (j0 + u < Nd) acc[u] = fmaf(w, __ldg(d + size_t(j0 + u) * V + c), acc[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
            if (lane == 0 && j0 + u < Nd) {
                srow[j0 + u] = acc[u];
                S[size_t(i) * Nd + j0 + u] = acc[u];
            }
        }
    }
    __syncthreads();
    if (la.mode < 0) return;
    // ---- (3) + (4)
    const float v = rank_loss_row<kGatherThreads>(la.mode, srow, la.teacher ? la.teacher + size_t(i) * Nd : nullptr,
                                                  la.dS ? la.dS + size_t(i) * Nd : nullptr, i, Nq, Nd, la.G, 1, la.invT, red);
    if (threadIdx.x == 0) la.rowloss[i] = v;
    finish_sum_last_block<kGatherThreads>(&g_ticket_gather, gridDim.x, la.rowloss, Nq, 1.f, la.loss, red);
}

// zero-fills S only when the dense fallback is going to accumulate split-K partials into it
__global__ void zero_if_dense_kernel(float* __restrict__ S, size_t n, const int* __restrict__ dense_flag) {
    if (*dense_flag == 0) return;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) S[i] = 0.f;
}
// dense fallback of the fused score+loss call: runs rank_loss rows only when the flag is set
__global__ void __launch_bounds__(256)
rank_loss_if_dense_kernel(int mode, const float* __restrict__ S, const float* __restrict__ teacher, int Nq, int C, int G,
                          int in_batch, float invT, const int* __restrict__ dense_flag, float* __restrict__ loss,
                          float* __restrict__ dS, float* __restrict__ rowloss) {
    __shared__ float red[8];
    if (*dense_flag == 0) return;
    const int i = blockIdx.x;
    const float v = rank_loss_row<256>(mode, S + size_t(i) * C, teacher ? teacher + size_t(i) * C : nullptr,
                                       dS ? dS + size_t(i) * C : nullptr, i, Nq, C, G, in_batch, invT, red);
    if (threadIdx.x == 0) rowloss[i] = v;
    finish_sum_last_block<256>(&g_ticket_rank_loss, gridDim.x, rowloss, Nq, 1.f, loss, red);
}

// d_d[j, :] (+)= gs * sum_i dS[i,j] * q[i,:] for rows [d_begin, d_end): the row is built in shared memory (zero-fill +
// shared-memory atomics over the query lists) and written to HBM once. gs = *gscale (device scalar) or 1.
__global__ void __launch_bounds__(kRowThreads, 1)
scores_docrow_bwd_kernel(const float* __restrict__ dS, const float* __restrict__ gscale, int Nq, int Nd, int V, int tpq,
                         QLists L, int d_begin, int d_end, int accumulate, float* __restrict__ d_d) {
    extern __shared__ float row_s[];
    if (*L.flag != 0) return;
    const float gs = gscale != nullptr ? __ldg(gscale) : 1.f;
    const int per = kRowThreads / tpq;
    const int sub = threadIdx.x % tpq;
    for (int j = d_begin + blockIdx.x; j < d_end; j += gridDim.x) {
        __syncthreads();
        for (int t = threadIdx.x; t < V; t += kRowThreads) row_s[t] = 0.f;
        __syncthreads();
        for (int q0 = 0; q0 < Nq; q0 += per) {   // all query chunks by the same block: one owner per output row
            const int qi = q0 + threadIdx.x / tpq;
            if (qi < Nq) {
                const float g = __ldg(dS + size_t(qi) * Nd + j) * gs;
                const int n = min(__ldg(L.nnz + qi), kQCap);
                if (g != 0.f)
                    for (int k = sub; k < n; k += tpq)
                        atomicAdd(&row_s[__ldg(L.cols + size_t(qi) * kQCap + k)], g * __ldg(L.vals + size_t(qi) * kQCap + k));
            }
        }
        __syncthreads();
        float* out = d_d + size_t(j) * V;
        for (int t = threadIdx.x; t < V; t += kRowThreads) out[t] = accumulate ? (out[t] + row_s[t]) : row_s[t];
    }
}

// ------------------------------------------------------------------------------------------ scores backward
// out[r, v] (+)= gs * sum_k coef(r,k) * in[k, v],  coef(r,k) = dS[r*sr + k*sk];  r in [r_begin, r_end), k in [0, K).
// 16 output rows per block, 2 columns per thread; coefficients staged in shared memory.
constexpr int kBR = 16, kBKc = 256;

template <int VECW>
__global__ void __launch_bounds__(256)
scores_bwd_kernel(const float* __restrict__ dS, const float* __restrict__ gscale, int sr, int sk,
                  const float* __restrict__ in, int K, int V, int r_begin, int r_end, int accumulate,
                  const int* __restrict__ dense_flag, float* __restrict__ out) {
    __shared__ float coef[kBR][kBKc];
    if (dense_flag != nullptr && *dense_flag == 0) return;  // the sparse-query kernel produced the rows
    const float gs = gscale != nullptr ? __ldg(gscale) : 1.f;
    const int r0 = r_begin + blockIdx.y * kBR;
    const int v = (blockIdx.x * 256 + threadIdx.x) * VECW;
    float acc[kBR][VECW];
#pragma unroll
    for (int r = 0; r < kBR; ++r)
#pragma unroll
        for (int e = 0; e < VECW; ++e) acc[r][e] = 0.f;
    for (int kc = 0; kc < K; kc += kBKc) {
        const int nk = min(kBKc, K - kc);
        __syncthreads();
        for (int t = threadIdx.x; t < kBR * nk; t += 256) {
            const int r = t / nk, k = t - r * nk;
            coef[r][k] = (r0 + r < r_end) ? __ldg(dS + size_t(r0 + r) * sr + size_t(kc + k) * sk) * gs : 0.f;
        }
        __syncthreads();
        if (v < V) {
#pragma unroll 4
            for (int k = 0; k < nk; ++k) {
                float x[VECW];
                const float* p = in + size_t(kc + k) * V + v;
                if (VECW == 2) {
                    const float2 t2 = __ldg(reinterpret_cast<const float2*>(p));
                    x[0] = t2.x;
                    x[VECW - 1] = t2.y;
                } else {
                    x[0] = __ldg(p);
                }
#pragma unroll
                for (int r = 0; r < kBR; ++r)
#pragma unroll
                    for (int e = 0; e < VECW; ++e) acc[r][e] = fmaf(coef[r][k], x[e], acc[r][e]);
            }
        }
    }
    if (v >= V) return;
#pragma unroll
    for (int r = 0; r < kBR; ++r) {
        if (r0 + r >= r_end) break;
        float* o = out + size_t(r0 + r) * V + v;
#pragma unroll
        for (int e = 0; e < VECW; ++e) o[e] = accumulate ? (o[e] + acc[r][e]) : acc[r][e];
    }
}

// own-docs backward: d_d[i*G+g, v] (+)= gs * dS[i,g] q[i,v]    grid (xb, rows)
__global__ void __launch_bounds__(256)
scores_group_bwd_d_kernel(const float* __restrict__ dS, const float* __restrict__ gscale, const float* __restrict__ q,
                          int G, int V, int d_begin, int accumulate, float* __restrict__ d_d) {
    const int row = d_begin + blockIdx.y;
    const int i = row / G;
    const float c = __ldg(dS + row) * (gscale != nullptr ? __ldg(gscale) : 1.f);  // dS is [Nq, G] row-major == index row
    const float* qi = q + size_t(i) * V;
    float* o = d_d + size_t(row) * V;
    for (int v = blockIdx.x * 256 + threadIdx.x; v < V; v += gridDim.x * 256) {
        const float gval = c * __ldg(qi + v);
        o[v] = accumulate ? (o[v] + gval) : gval;
    }
}
// own-docs backward: d_q[i, v] (+)= gs * sum_g dS[i,g] d[i*G+g, v]
__global__ void __launch_bounds__(256)
scores_group_bwd_q_kernel(const float* __restrict__ dS, const float* __restrict__ gscale, const float* __restrict__ d,
                          int G, int V, int q_begin, int accumulate, float* __restrict__ d_q) {
    const int i = q_begin + blockIdx.y;
    const float gs = gscale != nullptr ? __ldg(gscale) : 1.f;
    float* o = d_q + size_t(i) * V;
    for (int v = blockIdx.x * 256 + threadIdx.x; v < V; v += gridDim.x * 256) {
        float acc = 0.f;
        for (int g = 0; g < G; ++g) acc = fmaf(__ldg(dS + size_t(i) * G + g), __ldg(d + (size_t(i) * G + g) * V + v), acc);
        acc *= gs;
        o[v] = accumulate ? (o[v] + acc) : acc;
    }
}

// ------------------------------------------------------------------------------------------ CSR compaction
// Two streaming passes with one block per row (8-byte loads, 8 of them in flight per thread) around a one-block scan of
// the row counts: count the non-zeros of every row; ordered fill (each thread owns 4 consecutive columns of a 4096-column
// chunk, block scan of the per-thread counts), so columns stay ascending inside a row (= torch.nonzero order).
constexpr int kSeg = 1024;            // legacy segment width, still used by the scan kernel's interface (nseg = 1 here)
constexpr int kCsrThreads = 1024;

// loads elements [c, c+4) of a row (c % 4 == 0), zero beyond V / below first handled by the caller
__device__ __forceinline__ void load4(const float* __restrict__ row, int c, int V, bool vec2, float (&x)[4]) {
    if (vec2 && c + 4 <= V) {
        const float2 a = __ldg(reinterpret_cast<const float2*>(row + c));
        const float2 b = __ldg(reinterpret_cast<const float2*>(row + c + 2));
        x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) x[e] = (c + e < V) ? __ldg(row + c + e) : 0.f;
    }
}

__global__ void __launch_bounds__(kCsrThreads)
compact_count_kernel(const float* __restrict__ rep, int V, int first_col, int* __restrict__ rowcnt) {
    __shared__ float red[kCsrThreads / 32];
    const float* row = rep + size_t(blockIdx.x) * V;
    const bool vec2 = (reinterpret_cast<uintptr_t>(row) & 7) == 0;
    int cnt = 0;
    for (int c0 = 0; c0 < V; c0 += kCsrThreads * 4 * 2) {       // two chunks (8 eight-byte loads) in flight per thread
        float x[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) load4(row, c0 + u * kCsrThreads * 4 + threadIdx.x * 4, V, vec2, x[u]);
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int v = c0 + u * kCsrThreads * 4 + threadIdx.x * 4 + e;
                cnt += (v >= first_col && x[u][e] != 0.f) ? 1 : 0;
            }
    }
    const float total = block_sum<kCsrThreads>(float(cnt), red);   // exact: counts < 2^24
    if (threadIdx.x == 0) rowcnt[blockIdx.x] = int(total);
}

// exclusive scan of n counts -> offsets; row_ptr[b] = offset of the row's first segment; row_ptr[B] = total
__global__ void __launch_bounds__(1024)
compact_scan_kernel(const int* __restrict__ counts, int n, int nseg, int B, int* __restrict__ offsets,
                    int32_t* __restrict__ row_ptr) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int c = (i < n) ? counts[i] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = warp_tot[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_tot[lane] = wi - w;
        }
        __syncthreads();
        const int excl = carry_s + warp_tot[warp] + incl - c;
        if (i < n) {
            offsets[i] = excl;
            if (i % nseg == 0) row_ptr[i / nseg] = excl;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) row_ptr[B] = carry_s;
}

__global__ void __launch_bounds__(kCsrThreads)
compact_fill_kernel(const float* __restrict__ rep, int V, int first_col, const int* __restrict__ offsets,
                    int32_t* __restrict__ cols, float* __restrict__ vals, int capacity,
                    unsigned long long* __restrict__ df_count) {
    __shared__ int warp_tot[kCsrThreads / 32];
    const int b = blockIdx.x;
    const float* row = rep + size_t(b) * V;
    const bool vec2 = (reinterpret_cast<uintptr_t>(row) & 7) == 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int base = offsets[b];
    for (int c0 = 0; c0 < V; c0 += kCsrThreads * 4) {
        const int c = c0 + threadIdx.x * 4;
        float x[4];
        load4(row, c, V, vec2, x);
        int mine = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            // document frequency counts every column; only columns >= first_col are compacted
            if (df_count != nullptr && x[e] > 0.f) atomicAdd(df_count + c + e, 1ull);
            mine += (c + e >= first_col && x[e] != 0.f) ? 1 : 0;
        }
        int incl = mine;                                        // block-exclusive scan of the per-thread counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        __syncthreads();
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        const int wv = warp_tot[lane];                          // 32 warps: lane w holds warp w's total
        int winc = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        const int chunk_total = __shfl_sync(0xffffffffu, winc, 31);
        int off = base + __shfl_sync(0xffffffffu, winc - wv, warp) + incl - mine;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (c + e >= first_col && x[e] != 0.f) {
                if (off < capacity) {
                    cols[off] = c + e;
                    vals[off] = x[e];
                }
                ++off;
            }
        }
        base += chunk_total;
    }
}

// acc[i,:] (+)= scale * (S[i,:] - min) / (max - min + 1e-6)
__global__ void __launch_bounds__(256)
minmax_kernel(const float* __restrict__ S, int C, float scale, int accumulate, float* __restrict__ acc) {
    __shared__ float red[8];
    const float* s = S + size_t(blockIdx.x) * C;
    float* a = acc + size_t(blockIdx.x) * C;
    float mx = -CUDART_INF_F, mn = CUDART_INF_F;
    for (int j = threadIdx.x; j < C; j += 256) {
        mx = fmaxf(mx, s[j]);
        mn = fminf(mn, s[j]);
    }
    mx = block_max<256>(mx, red);
    mn = block_min<256>(mn, red);
    const float den = (mx - mn) + 1e-6f;
    for (int j = threadIdx.x; j < C; j += 256) {
        const float v = (s[j] - mn) / den * scale;
        a[j] = accumulate ? (a[j] + v) : v;
    }
}

int pick_ksplit(int base_blocks, int V, int quantum, int* kchunk) {
    int ks = (2 * num_sms() + base_blocks - 1) / base_blocks;
    if (ks < 1) ks = 1;
    int kc = (V + ks - 1) / ks;
    kc = int(align_up(size_t(kc), size_t(quantum)));
    ks = (V + kc - 1) / kc;
    *kchunk = kc;
    return ks;
}

struct RowPlan {
    int n_parts, part_len;
    size_t smem;
};
// a row is cut into the smallest number (>= 2 for long rows: finer pipelining) of equal parts such that kRowStages
// stages fit the shared-memory budget
RowPlan row_plan(int V) {
    RowPlan p;
    p.n_parts = V >= 8192 ? 2 : 1;
    while (true) {
        p.part_len = int(align_up(size_t((V + p.n_parts - 1) / p.n_parts), 4));
        p.smem = size_t(kRowStages) * size_t(p.part_len + 8) * sizeof(float);
        if (p.smem <= size_t(kRowSmemBudget)) break;
        ++p.n_parts;
    }
    return p;
}

// Very sparse queries (caller-promised bound): gathering Nq*bound 32-byte sectors per document beats streaming its
// V*4 bytes, and a block per query needs no grid-wide barrier for the loss -> one launch.
bool use_gather_kernel(int Nq, int Nd, int V, int q_nnz_bound) {
    return q_nnz_bound > 0 && q_nnz_bound <= kGatherCap && Nd <= kGatherMaxNd &&
           (long long)Nq * q_nnz_bound * 8 <= (long long)V;
}
int launch_gather(const float* q, const float* d, int Nq, int Nd, int V, QLists L, float* S, const LossArgs& la,
                  cudaStream_t stream) {
    score_gather_kernel<<<Nq, kGatherThreads, size_t(Nd) * sizeof(float), stream>>>(q, d, Nq, Nd, V, L, S, la);
    SB200_CHECK_LAUNCH("score_gather_kernel");
    return SB200_OK;
}

int launch_q_compact(const float* q, int Nq, int V, QLists L, cudaStream_t stream) {
    q_compact_kernel<<<Nq, kCompactThreads, 0, stream>>>(q, V, L);
    SB200_CHECK_LAUNCH("q_compact_kernel");
    return SB200_OK;
}

// the persistent row kernel; fused != 0 -> cooperative launch with the loss phase
int launch_docrow(const float* d, int Nq, int Nd, int V, QLists L, float* S, const LossArgs& la, bool fused,
                  cudaStream_t stream) {
    const RowPlan plan = row_plan(V);
    int grid = num_sms();
    if (grid > Nd) grid = Nd;
    if (!device_flag_test_and_set(6)) {
        SB200_CUDA(cudaFuncSetAttribute(scores_docrow_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kRowSmemBudget));
        SB200_CUDA(cudaFuncSetAttribute(scores_docrow_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kRowSmemBudget));
    }
    int n_parts = plan.n_parts, part_len = plan.part_len;
    if (fused) {
        void* args[] = {(void*)&d, (void*)&Nq, (void*)&Nd, (void*)&V, (void*)&n_parts, (void*)&part_len, (void*)&L,
                        (void*)&S, (void*)&la};
        SB200_CUDA(cudaLaunchCooperativeKernel((const void*)scores_docrow_kernel<true>, dim3(grid), dim3(kRowThreads), args,
                                               plan.smem, stream));
    } else {
        scores_docrow_kernel<false><<<grid, kRowThreads, plan.smem, stream>>>(d, Nq, Nd, V, n_parts, part_len, L, S, la);
    }
    SB200_CHECK_LAUNCH("scores_docrow_kernel");
    return SB200_OK;
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" size_t sb200_scores_workspace_bytes(int Nq, int Nd, int V, int in_batch) {
    (void)Nd; (void)V;
    if (Nq <= 0) return 0;
    // in-batch: thresholded query lists (+ dispatch flag), reused by sb200_scores_bwd; both modes: Nq row losses
    return (in_batch ? qlists_bytes(Nq) : 0) + align_up_sz(size_t(Nq) * sizeof(float), 256);
}

static int row_kernel_grid(int rows, int chunks) {
    int g = num_sms() / chunks;  // one row (V*4 B of shared memory) per SM at a time
    if (g < 1) g = 1;
    return g < rows ? g : rows;
}
static size_t row_slab_bytes(int V) { return size_t(V + 4) * sizeof(float); }
// threads that share one query (power of two <= 32): as many as the block allows, so short lists stay in registers
static int threads_per_query(int Nq) {
    int tpq = 32;
    while (tpq > 2 && kRowThreads / tpq < Nq) tpq >>= 1;
    return tpq;
}

// dense fp32 tiles: the only path without a workspace, the fallback (device-side flag) with one
static int launch_dense_scores(const float* q, const float* d, int Nq, int Nd, int V, const int* dense_flag, float* S,
                               cudaStream_t stream) {
    const int tj = (Nd + kST - 1) / kST, ti = (Nq + kST - 1) / kST;
    SB200_REQUIRE(ti <= 65535, "scores_fwd: Nq too large");
    int kchunk;
    const int ks = pick_ksplit(tj * ti, V, kSK, &kchunk);
    if (ks > 1) {
        if (dense_flag != nullptr) {
            zero_if_dense_kernel<<<2 * num_sms(), 256, 0, stream>>>(S, size_t(Nq) * Nd, dense_flag);
            SB200_CHECK_LAUNCH("zero_if_dense_kernel");
        } else {
            SB200_CUDA(cudaMemsetAsync(S, 0, size_t(Nq) * Nd * sizeof(float), stream));
        }
    }
    const bool mma_ok = (V % 2 == 0) && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(d)) & 7) == 0;
    if (mma_ok) scores_tile_mma_kernel<<<dim3(tj, ti, ks), 256, 0, stream>>>(q, d, Nq, Nd, V, kchunk, dense_flag, S);
    else scores_tile_kernel<<<dim3(tj, ti, ks), 256, 0, stream>>>(q, d, Nq, Nd, V, kchunk, dense_flag, S);
    SB200_CHECK_LAUNCH("scores_tile_kernel");
    return SB200_OK;
}

static int launch_group(const float* q, const float* d, int Nq, int G, int V, int mode, const float* teacher, float invT,
                        float* S, float* dS, float* rowloss, float* loss, cudaStream_t stream) {
    SB200_REQUIRE(Nq <= 65535, "scores: Nq too large");
    SB200_REQUIRE(G <= kGroupMaxG, "scores (own docs): G=%d exceeds %d", G, kGroupMaxG);
    int ks = (2 * num_sms() + Nq - 1) / Nq;   // CTAs per query: fill the machine, at most a portable cluster
    if (ks > 8) ks = 8;
    while (ks & (ks - 1)) --ks;
    if (ks < 1) ks = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(ks), unsigned(Nq));
    cfg.blockDim = dim3(kGroupThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = unsigned(ks);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SB200_CUDA(cudaLaunchKernelEx(&cfg, score_group_kernel, q, d, Nq, G, V, mode, teacher, invT, S, dS, rowloss, loss));
    SB200_CHECK_LAUNCH("score_group_kernel");
    return SB200_OK;
}

static float* rowloss_of(void* workspace, int Nq, int in_batch) {
    return reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + (in_batch ? qlists_bytes(Nq) : 0));
}

extern "C" int sb200_scores_fwd(const float* q, const float* d, int Nq, int Nd, int V, int in_batch, int q_nnz_bound,
                                float* S, void* workspace, size_t workspace_bytes, sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(q && d && S, "scores_fwd: null pointer");
    SB200_REQUIRE(Nq >= 1 && Nd >= 1 && V >= 1, "scores_fwd: bad shape");
    if (in_batch) {
        const bool sparse_ok = workspace != nullptr && workspace_bytes >= qlists_bytes(Nq) && Nq <= 65535;
        const int* dense_flag = nullptr;
        if (sparse_ok) {
            QLists L = qlists_carve(workspace, Nq);
            int rc = launch_q_compact(q, Nq, V, L, stream);
            if (rc != SB200_OK) return rc;
            LossArgs la = {};
            la.mode = -1;
            rc = launch_docrow(d, Nq, Nd, V, L, S, la, false, stream);
            if (rc != SB200_OK) return rc;
            dense_flag = L.flag;
            // a promised bound on the non-zeros per query row: the sparse kernel always does the work
            if (q_nnz_bound > 0 && q_nnz_bound <= kQCap && Nq <= kRowMaxQ) return SB200_OK;
        }
        return launch_dense_scores(q, d, Nq, Nd, V, dense_flag, S, stream);
    }
    SB200_REQUIRE(Nd % Nq == 0, "scores_fwd: Nd=%d is not a multiple of Nq=%d", Nd, Nq);
    return launch_group(q, d, Nq, Nd / Nq, V, -1, nullptr, 1.f, S, nullptr, nullptr, nullptr, stream);
}

extern "C" int sb200_score_loss_fwd(int mode, const float* q, const float* d, const float* teacher, int Nq, int Nd, int V,
                                    int G, int in_batch, float temperature, int q_nnz_bound, float* S, float* loss,
                                    float* dS, void* workspace, size_t workspace_bytes, sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(q && d && S && loss, "score_loss_fwd: null pointer");
    SB200_REQUIRE(Nq >= 1 && Nd >= 1 && V >= 1 && G >= 1, "score_loss_fwd: bad shape");
    SB200_REQUIRE(mode >= SB200_LOSS_INFONCE && mode <= SB200_LOSS_MARGINMSE, "score_loss_fwd: bad mode %d", mode);
    SB200_REQUIRE(mode == SB200_LOSS_INFONCE || teacher != nullptr, "score_loss_fwd: teacher scores required");
    SB200_REQUIRE(temperature > 0.f, "score_loss_fwd: temperature must be positive");
    SB200_REQUIRE(Nd == Nq * G, "score_loss_fwd: Nd=%d != Nq*G=%d", Nd, Nq * G);
    const int C = in_batch ? Nd : G;
    SB200_REQUIRE(mode != SB200_LOSS_MARGINMSE || C >= 2, "score_loss_fwd: marginmse needs >= 2 columns");
    const size_t need = sb200_scores_workspace_bytes(Nq, Nd, V, in_batch);
    if (workspace == nullptr || workspace_bytes < need)
        return fail(SB200_ERR_WORKSPACE, "score_loss_fwd: workspace %zu < %zu", workspace_bytes, need);
    float* rowloss = rowloss_of(workspace, Nq, in_batch);
    const float invT = 1.f / temperature;
    if (!in_batch)
        return launch_group(q, d, Nq, G, V, mode, teacher, invT, S, dS, rowloss, loss, stream);
    SB200_REQUIRE(Nq <= 65535, "score_loss_fwd: Nq too large");
    QLists L = qlists_carve(workspace, Nq);
    LossArgs lg;
    lg.mode = mode; lg.teacher = teacher; lg.G = G; lg.invT = invT; lg.loss = loss; lg.dS = dS; lg.rowloss = rowloss;
    if (use_gather_kernel(Nq, Nd, V, q_nnz_bound)) return launch_gather(q, d, Nq, Nd, V, L, S, lg, stream);
    int rc = launch_q_compact(q, Nq, V, L, stream);
    if (rc != SB200_OK) return rc;
    LossArgs la;
    la.mode = mode;
    la.teacher = teacher;
    la.G = G;
    la.invT = invT;
    la.loss = loss;
    la.dS = dS;
    la.rowloss = rowloss;
    rc = launch_docrow(d, Nq, Nd, V, L, S, la, true, stream);
    if (rc != SB200_OK) return rc;
    // Caller-guaranteed bound on the non-zeros per query row (inf-free queries: the token count): the sparse kernel
    // always does the work and nothing else is launched. Otherwise the dense kernels follow; they exit at once unless
    // the device-side flag says a query row did not fit the lists.
    if (q_nnz_bound > 0 && q_nnz_bound <= kQCap && Nq <= kRowMaxQ) return SB200_OK;
    rc = launch_dense_scores(q, d, Nq, Nd, V, L.flag, S, stream);
    if (rc != SB200_OK) return rc;
    rank_loss_if_dense_kernel<<<Nq, 256, 0, stream>>>(mode, S, teacher, Nq, C, G, 1, invT, L.flag, loss, dS, rowloss);
    SB200_CHECK_LAUNCH("rank_loss_if_dense_kernel");
    return SB200_OK;
}

extern "C" int sb200_scores_bwd(const float* dS, const float* gscale, const float* q, const float* d, int Nq, int Nd,
                                int V, int in_batch, int q_begin, int q_end, int d_begin, int d_end, int accumulate,
                                float* d_q, float* d_d, const void* fwd_workspace, size_t workspace_bytes,
                                sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(dS && q && d, "scores_bwd: null pointer");
    SB200_REQUIRE(Nq >= 1 && Nd >= 1 && V >= 1, "scores_bwd: bad shape");
    SB200_REQUIRE(0 <= q_begin && q_begin <= q_end && q_end <= Nq && 0 <= d_begin && d_begin <= d_end && d_end <= Nd,
                  "scores_bwd: bad row ranges");
    const bool vec2 = (V % 2 == 0) && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(d) |
                                        reinterpret_cast<uintptr_t>(d_q) | reinterpret_cast<uintptr_t>(d_d)) & 7) == 0;
    if (in_batch) {
        const int xb = vec2 ? (V / 2 + 255) / 256 : (V + 255) / 256;
        if (d_d != nullptr && d_end > d_begin) {
            const size_t row_smem = row_slab_bytes(V);
            const int* dense_flag = nullptr;
            if (fwd_workspace != nullptr && workspace_bytes >= qlists_bytes(Nq) && row_smem <= size_t(kRowSmemBudget)) {
                // query lists built by the forward call: one shared-memory row per document, written to HBM once
                QLists L = qlists_carve(const_cast<void*>(fwd_workspace), Nq);
                if (!device_flag_test_and_set(7))
                    SB200_CUDA(cudaFuncSetAttribute(scores_docrow_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    kRowSmemBudget));
                scores_docrow_bwd_kernel<<<row_kernel_grid(d_end - d_begin, 1), kRowThreads, row_smem, stream>>>(
                    dS, gscale, Nq, Nd, V, threads_per_query(Nq), L, d_begin, d_end, accumulate, d_d);
                SB200_CHECK_LAUNCH("scores_docrow_bwd_kernel");
                dense_flag = L.flag;
            }
            dim3 grid(xb, (d_end - d_begin + kBR - 1) / kBR);
            // out row r = doc j, k = query i: coef = dS[i*Nd + j]
            if (vec2) scores_bwd_kernel<2><<<grid, 256, 0, stream>>>(dS, gscale, 1, Nd, q, Nq, V, d_begin, d_end, accumulate, dense_flag, d_d);
            else scores_bwd_kernel<1><<<grid, 256, 0, stream>>>(dS, gscale, 1, Nd, q, Nq, V, d_begin, d_end, accumulate, dense_flag, d_d);
            SB200_CHECK_LAUNCH("scores_bwd_kernel(d_d)");
        }
        if (d_q != nullptr && q_end > q_begin) {
            dim3 grid(xb, (q_end - q_begin + kBR - 1) / kBR);
            if (vec2) scores_bwd_kernel<2><<<grid, 256, 0, stream>>>(dS, gscale, Nd, 1, d, Nd, V, q_begin, q_end, accumulate, nullptr, d_q);
            else scores_bwd_kernel<1><<<grid, 256, 0, stream>>>(dS, gscale, Nd, 1, d, Nd, V, q_begin, q_end, accumulate, nullptr, d_q);
            SB200_CHECK_LAUNCH("scores_bwd_kernel(d_q)");
        }
    } else {
        SB200_REQUIRE(Nd % Nq == 0, "scores_bwd: Nd=%d is not a multiple of Nq=%d", Nd, Nq);
        const int G = Nd / Nq;
        int xb = (V + 255) / 256;
        if (xb > 32) xb = 32;
        if (d_d != nullptr && d_end > d_begin) {
            SB200_REQUIRE(d_end - d_begin <= 65535, "scores_bwd: too many rows");
            scores_group_bwd_d_kernel<<<dim3(xb, d_end - d_begin), 256, 0, stream>>>(dS, gscale, q, G, V, d_begin, accumulate, d_d);
            SB200_CHECK_LAUNCH("scores_group_bwd_d_kernel");
        }
        if (d_q != nullptr && q_end > q_begin) {
            SB200_REQUIRE(q_end - q_begin <= 65535, "scores_bwd: too many rows");
            scores_group_bwd_q_kernel<<<dim3(xb, q_end - q_begin), 256, 0, stream>>>(dS, gscale, d, G, V, q_begin, accumulate, d_q);
            SB200_CHECK_LAUNCH("scores_group_bwd_q_kernel");
        }
    }
    return SB200_OK;
}

extern "C" size_t sb200_rank_loss_workspace_bytes(int Nq) {
    return Nq > 0 ? align_up_sz(size_t(Nq) * sizeof(float), 256) : 0;
}

extern "C" int sb200_rank_loss(int mode, const float* S, const float* teacher, int Nq, int C, int G, int in_batch,
                               float temperature, float* loss, float* dS, void* workspace, size_t workspace_bytes,
                               sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(S && loss, "rank_loss: null pointer");
    SB200_REQUIRE(mode >= SB200_LOSS_INFONCE && mode <= SB200_LOSS_MARGINMSE, "rank_loss: bad mode %d", mode);
    SB200_REQUIRE(Nq >= 1 && C >= 1 && G >= 1, "rank_loss: bad shape");
    SB200_REQUIRE(mode == SB200_LOSS_INFONCE || teacher != nullptr, "rank_loss: teacher scores required");
    SB200_REQUIRE(mode != SB200_LOSS_MARGINMSE || C >= 2, "rank_loss: marginmse needs >= 2 columns");
    SB200_REQUIRE(temperature > 0.f, "rank_loss: temperature must be positive");
    if (mode == SB200_LOSS_INFONCE) {
        if (in_batch) SB200_REQUIRE(C == Nq * G, "rank_loss: in-batch infonce expects C == Nq*G");
        else SB200_REQUIRE(C == G, "rank_loss: infonce expects C == G");
    }
    if (workspace == nullptr || workspace_bytes < sb200_rank_loss_workspace_bytes(Nq))
        return fail(SB200_ERR_WORKSPACE, "rank_loss: workspace too small");
    rank_loss_kernel<<<Nq, 256, 0, stream>>>(mode, S, teacher, Nq, C, G, in_batch, 1.f / temperature, loss, dS,
                                             static_cast<float*>(workspace));
    SB200_CHECK_LAUNCH("rank_loss_kernel");
    return SB200_OK;
}

extern "C" size_t sb200_compact_workspace_bytes(int B, int V) {
    if (B <= 0 || V <= 0) return 0;
    return 2 * align_up(size_t(B) * sizeof(int), 256);     // row counts, row offsets
}

extern "C" int sb200_compact_rows(const float* rep, int B, int V, int first_col, int32_t* row_ptr, int32_t* cols,
                                  float* vals, int capacity, int64_t* df_count, void* workspace, size_t workspace_bytes,
                                  sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(rep && row_ptr && cols && vals, "compact_rows: null pointer");
    SB200_REQUIRE(B >= 1 && V >= 1 && first_col >= 0 && capacity >= 0, "compact_rows: bad shape");
    if (workspace == nullptr || workspace_bytes < sb200_compact_workspace_bytes(B, V))
        return fail(SB200_ERR_WORKSPACE, "compact_rows: workspace too small");
    int* rowcnt = static_cast<int*>(workspace);
    int* rowoff = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + align_up(size_t(B) * sizeof(int), 256));
    compact_count_kernel<<<B, kCsrThreads, 0, stream>>>(rep, V, first_col, rowcnt);
    SB200_CHECK_LAUNCH("compact_count_kernel");
    compact_scan_kernel<<<1, 1024, 0, stream>>>(rowcnt, B, 1, B, rowoff, row_ptr);
    SB200_CHECK_LAUNCH("compact_scan_kernel");
    compact_fill_kernel<<<B, kCsrThreads, 0, stream>>>(rep, V, first_col, rowoff, cols, vals, capacity,
                                                       reinterpret_cast<unsigned long long*>(df_count));
    SB200_CHECK_LAUNCH("compact_fill_kernel");
    return SB200_OK;
}

extern "C" int sb200_minmax_accumulate(const float* S, int Nq, int C, float scale, int accumulate, float* acc,
                                       sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(S && acc && Nq >= 1 && C >= 1, "minmax_accumulate: bad arguments");
    minmax_kernel<<<Nq, 256, 0, stream>>>(S, C, scale, accumulate, acc);
    SB200_CHECK_LAUNCH("minmax_kernel");
    return SB200_OK;
}
