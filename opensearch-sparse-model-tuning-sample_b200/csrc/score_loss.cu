// In-batch query x doc scoring and the ranking losses of scripts/train/loss.py, plus the encode-output
// compaction (sparse_encoders.py:137-150, 178-179) and the teacher min-max normalisation
// (bi_encoder_wrapper.py:133-138). fp32 CUDA-core arithmetic throughout (the reference runs these GEMMs in
// fp32 with TF32 off); the score kernels are HBM/L2-bound: (Nq+Nd)*V*4 bytes against Nq*Nd*V*2 flop.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "common.h"
#include "ptx.cuh"

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

// ------------------------------------------------------------------------------------------ scores, in-batch
// S[i,j] += sum_{k in split} q[i,k] d[j,k].  64x64 output tile per block, 4x4 per thread, K staged 32 at a time
// through shared memory (transposed, stride 65 -> conflict-free), split-K over blockIdx.z merged with atomics.
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

// scores, own docs only: S[i,g] = q[i,:] . d[i*G+g,:]   grid (ksplit, Nq)
__global__ void __launch_bounds__(256)
scores_group_kernel(const float* __restrict__ q, const float* __restrict__ d, int G, int V, int kchunk,
                    float* __restrict__ S) {
    __shared__ float red[8];
    const int i = blockIdx.y;
    const int k_begin = blockIdx.x * kchunk, k_end = min(V, k_begin + kchunk);
    const float* qi = q + size_t(i) * V;
    for (int g0 = 0; g0 < G; g0 += 8) {
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        for (int k = k_begin + threadIdx.x; k < k_end; k += 256) {
            const float qv = __ldg(qi + k);
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (g0 + e < G) acc[e] = fmaf(qv, __ldg(d + (size_t(i) * G + g0 + e) * V + k), acc[e]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (g0 + e < G) {  // block-uniform
                const float s = block_sum<256>(acc[e], red);
                if (threadIdx.x == 0) {
                    if (gridDim.x == 1) S[size_t(i) * G + g0 + e] = s;
                    else atomicAdd(S + size_t(i) * G + g0 + e, s);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ sparse-query path
// Queries are very sparse (inf-free: at most Lq token ids; learned: tens to hundreds of entries once trained), so
// q.d^T = sum over the query's non-zeros of val * d[j, col]. The query rows are thresholded (!= 0) into (col, val)
// lists once; then each DOCUMENT row is staged in shared memory (V*4 B = 122 KB) exactly once and every query's list
// is gathered from it: HBM traffic = one pass over d, S[i,j] written once, no atomics (deterministic).
// Dispatch is on the device: if any query row has more than kQCap non-zeros the flag is set, the sparse kernels exit
// and the dense tile kernel (which exits when the flag is clear) does the work instead -- no host synchronisation.
constexpr int kQCap = 512;
constexpr int kRowThreads = 512;

struct QLists {
    int* flag;      // [1]  1 = some row overflowed -> dense path
    int* nnz;       // [Nq]
    int* cols;      // [Nq][kQCap]
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

// grid (column slabs, Nq): slots are claimed with one atomic per non-zero on the row's counter (zeroed by the caller);
// the order of a list is irrelevant. nnz[i] ends up as the true count (may exceed the capacity: then the flag is set).
__global__ void __launch_bounds__(256)
q_compact_kernel(const float* __restrict__ q, int V, int cap, QLists L) {
    const int i = blockIdx.y;
    const float* row = q + size_t(i) * V;
    for (int v = blockIdx.x * 256 + threadIdx.x; v < V; v += gridDim.x * 256) {
        const float x = __ldg(row + v);
        if (x != 0.f) {
            const int slot = atomicAdd(L.nnz + i, 1);
            if (slot < kQCap) {
                L.cols[size_t(i) * kQCap + slot] = v;
                L.vals[size_t(i) * kQCap + slot] = x;
            }
            if (slot == cap) atomicOr(L.flag, 1);  // the row does not fit the register budget of the row kernels
        }
    }
}

// Asynchronous global->shared copy of src[0, n) (cp.async, no registers held): every thread issues all its
// requests before anyone waits, so a whole slab is in flight per CTA.
__device__ __forceinline__ void stage_async(float* __restrict__ dst, const float* __restrict__ src, int n) {
    const uint32_t d0 = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
    if ((reinterpret_cast<uintptr_t>(src) & 7) == 0) {
        const int n2 = n >> 1;
        for (int t = threadIdx.x; t < n2; t += kRowThreads)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0 + t * 8), "l"(src + 2 * t) : "memory");
        if ((n & 1) && threadIdx.x == 0)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + (n - 1) * 4), "l"(src + n - 1) : "memory");
    } else {
        for (int t = threadIdx.x; t < n; t += kRowThreads)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + t * 4), "l"(src + t) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// grid (row slots, query chunks of kRowThreads/tpq queries). Each thread keeps up to kEPT (column, value) entries of ONE
// query in registers for the whole kernel (tpq threads share a query); a block walks document rows: the row is
// staged in shared memory with cp.async (whole row in flight), every thread gathers its entries from it, the tpq
// partial sums are shuffled together and S[i,j] is written once -- no atomics, deterministic, d is read from HBM once
// per query chunk. Rows longer than the register budget (nnz > kEPT*tpq) raise the device-side flag instead.
constexpr int kEPT = 32;

__global__ void __launch_bounds__(kRowThreads, 1)
scores_docrow_kernel(const float* __restrict__ d, int Nq, int Nd, int V, int tpq, QLists L, float* __restrict__ S) {
    extern __shared__ __align__(128) float row_s[];
    if (*L.flag != 0) return;  // dense path handles it
    const int qi = blockIdx.y * (kRowThreads / tpq) + threadIdx.x / tpq;
    const int sub = threadIdx.x % tpq;
    int col[kEPT];
    float val[kEPT];
    const int n = (qi < Nq) ? min(__ldg(L.nnz + qi), kQCap) : 0;
#pragma unroll
    for (int e = 0; e < kEPT; ++e) {
        const int k = sub + e * tpq;
        const bool ok = k < n;
        col[e] = ok ? __ldg(L.cols + size_t(qi) * kQCap + k) : 0;
        val[e] = ok ? __ldg(L.vals + size_t(qi) * kQCap + k) : 0.f;
    }
    // rows are pulled with ONE bulk async copy each (TMA engine, completion on an mbarrier); the few floats before
    // the first / after the last 16-byte boundary are moved by ordinary loads. `shift` keeps source and destination
    // in the same 16-byte phase.
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    for (int j = blockIdx.x; j < Nd; j += gridDim.x) {
        const float* src = d + size_t(j) * V;
        const int head = min(V, int(((16 - (reinterpret_cast<uintptr_t>(src) & 15)) & 15) >> 2));
        const int shift = (4 - head) & 3;
        const int nbulk = ((V - head) >> 2) << 2;
        __syncthreads();  // everyone is done gathering from the previous row
        if (threadIdx.x == 0 && nbulk > 0) {
            mbar_arrive_expect_tx(&bar, uint32_t(nbulk) * 4u);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(row_s + head + shift)), "l"(src + head), "r"(uint32_t(nbulk) * 4u), "r"(smem_u32(&bar))
                         : "memory");
        }
        if (int(threadIdx.x) < head) row_s[threadIdx.x + shift] = __ldg(src + threadIdx.x);
        for (int t = head + nbulk + threadIdx.x; t < V; t += kRowThreads) row_s[t + shift] = __ldg(src + t);
        if (nbulk > 0) mbar_wait(&bar, phase);
        phase ^= 1;
        __syncthreads();
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < kEPT; ++e) acc = fmaf(val[e], row_s[col[e] + shift], acc);
        for (int o = tpq >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (sub == 0 && qi < Nq) S[size_t(qi) * Nd + j] = acc;
    }
}

// zero-fills S only when the dense fallback is going to accumulate split-K partials into it
__global__ void zero_if_dense_kernel(float* __restrict__ S, size_t n, const int* __restrict__ dense_flag) {
    if (*dense_flag == 0) return;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) S[i] = 0.f;
}

// d_d[j, :] (+)= sum_i dS[i,j] * q[i,:] for rows [d_begin, d_end): same register-resident lists; the row is built in
// shared memory (zero-fill + shared-memory atomics) and written to HBM once. Query chunks beyond the first accumulate.
__global__ void __launch_bounds__(kRowThreads, 1)
scores_docrow_bwd_kernel(const float* __restrict__ dS, int Nq, int Nd, int V, int tpq, QLists L, int d_begin, int d_end,
                         int accumulate, float* __restrict__ d_d) {
    extern __shared__ float row_s[];
    if (*L.flag != 0) return;
    const int per = kRowThreads / tpq;
    const int sub = threadIdx.x % tpq;
    for (int j = d_begin + blockIdx.x; j < d_end; j += gridDim.x) {
        __syncthreads();
        for (int t = threadIdx.x; t < V; t += kRowThreads) row_s[t] = 0.f;
        __syncthreads();
        for (int q0 = 0; q0 < Nq; q0 += per) {   // all query chunks by the same block: one owner per output row
            const int qi = q0 + threadIdx.x / tpq;
            if (qi < Nq) {
                const float g = __ldg(dS + size_t(qi) * Nd + j);
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
// out[r, v] (+)= sum_k coef(r,k) * in[k, v],  coef(r,k) = dS[r*sr + k*sk];  r in [r_begin, r_end), k in [0, K).
// 16 output rows per block, 2 columns per thread; coefficients staged in shared memory.
constexpr int kBR = 16, kBKc = 256;

template <int VECW>
__global__ void __launch_bounds__(256)
scores_bwd_kernel(const float* __restrict__ dS, int sr, int sk, const float* __restrict__ in, int K, int V, int r_begin,
                  int r_end, int accumulate, const int* __restrict__ dense_flag, float* __restrict__ out) {
    __shared__ float coef[kBR][kBKc];
    if (dense_flag != nullptr && *dense_flag == 0) return;  // the sparse-query kernel produced the rows
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
            coef[r][k] = (r0 + r < r_end) ? __ldg(dS + size_t(r0 + r) * sr + size_t(kc + k) * sk) : 0.f;
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

// own-docs backward: d_d[i*G+g, v] (+)= dS[i,g] q[i,v]    grid (xb, rows)
__global__ void __launch_bounds__(256)
scores_group_bwd_d_kernel(const float* __restrict__ dS, const float* __restrict__ q, int G, int V, int d_begin,
                          int accumulate, float* __restrict__ d_d) {
    const int row = d_begin + blockIdx.y;
    const int i = row / G;
    const float c = __ldg(dS + row);  // dS is [Nq, G] row-major == index row
    const float* qi = q + size_t(i) * V;
    float* o = d_d + size_t(row) * V;
    for (int v = blockIdx.x * 256 + threadIdx.x; v < V; v += gridDim.x * 256) {
        const float gval = c * __ldg(qi + v);
        o[v] = accumulate ? (o[v] + gval) : gval;
    }
}
// own-docs backward: d_q[i, v] (+)= sum_g dS[i,g] d[i*G+g, v]
__global__ void __launch_bounds__(256)
scores_group_bwd_q_kernel(const float* __restrict__ dS, const float* __restrict__ d, int G, int V, int q_begin,
                          int accumulate, float* __restrict__ d_q) {
    const int i = q_begin + blockIdx.y;
    float* o = d_q + size_t(i) * V;
    for (int v = blockIdx.x * 256 + threadIdx.x; v < V; v += gridDim.x * 256) {
        float acc = 0.f;
        for (int g = 0; g < G; ++g) acc = fmaf(__ldg(dS + size_t(i) * G + g), __ldg(d + (size_t(i) * G + g) * V + v), acc);
        o[v] = accumulate ? (o[v] + acc) : acc;
    }
}

// ------------------------------------------------------------------------------------------ ranking losses
// One block per query row. loss is accumulated with one atomicAdd per row (already scaled by the batch mean).
__global__ void __launch_bounds__(256)
rank_loss_kernel(int mode, const float* __restrict__ S, const float* __restrict__ teacher, int Nq, int C, int G,
                 int in_batch, float invT, float* __restrict__ loss, float* __restrict__ dS) {
    __shared__ float red[8];
    const int i = blockIdx.x;
    const float* s = S + size_t(i) * C;
    float* g = dS != nullptr ? dS + size_t(i) * C : nullptr;
    const float invNq = 1.f / float(Nq);

    if (mode == SB200_LOSS_INFONCE) {
        // selected columns: own positive + every hard negative (loss.py:90-101); own docs only when !in_batch
        const int pos = in_batch ? i * G : 0;
        auto selected = [&](int j) { return !in_batch || j == pos || (j % G) != 0; };
        float m = -CUDART_INF_F;
        for (int j = threadIdx.x; j < C; j += 256)
            if (selected(j)) m = fmaxf(m, s[j]);
        m = block_max<256>(m, red);
        float z = 0.f;
        for (int j = threadIdx.x; j < C; j += 256)
            if (selected(j)) z += expf(s[j] - m);
        z = block_sum<256>(z, red);
        const float lse = m + logf(z);
        if (threadIdx.x == 0) atomicAdd(loss, (lse - s[pos]) * invNq);
        if (g != nullptr) {
            for (int j = threadIdx.x; j < C; j += 256) {
                float v = 0.f;
                if (selected(j)) v = (expf(s[j] - lse) - (j == pos ? 1.f : 0.f)) * invNq;
                g[j] = v;
            }
        }
    } else if (mode == SB200_LOSS_KLDIV) {
        const float* t = teacher + size_t(i) * C;
        float ms = -CUDART_INF_F, mt = -CUDART_INF_F;
        for (int j = threadIdx.x; j < C; j += 256) {
            ms = fmaxf(ms, s[j] * invT);
            mt = fmaxf(mt, t[j] * invT);
        }
        ms = block_max<256>(ms, red);
        mt = block_max<256>(mt, red);
        float zs = 0.f, zt = 0.f;
        for (int j = threadIdx.x; j < C; j += 256) {
            zs += expf(s[j] * invT - ms);
            zt += expf(t[j] * invT - mt);
        }
        zs = block_sum<256>(zs, red);
        zt = block_sum<256>(zt, red);
        const float lzs = logf(zs), lzt = logf(zt);
        float acc = 0.f;
        for (int j = threadIdx.x; j < C; j += 256) {
            const float lps = s[j] * invT - ms - lzs;
            const float lpt = t[j] * invT - mt - lzt;
            const float pt = expf(lpt);
            if (pt > 0.f) acc += pt * (lpt - lps);  // xlogy(t,t) - t*input
            if (g != nullptr) g[j] = (expf(lps) - pt) * invT * invNq;
        }
        acc = block_sum<256>(acc, red);
        if (threadIdx.x == 0) atomicAdd(loss, acc * invNq);
    } else {  // margin MSE: margins against column 0 (loss.py:52-55)
        const float* t = teacher + size_t(i) * C;
        const float s0 = s[0] * invT, t0 = t[0] * invT;
        const float norm = 1.f / (float(Nq) * float(C - 1));
        float acc = 0.f, g0 = 0.f;
        for (int j = 1 + threadIdx.x; j < C; j += 256) {
            const float diff = (s0 - s[j] * invT) - (t0 - t[j] * invT);
            acc = fmaf(diff, diff, acc);
            const float gj = 2.f * diff * norm * invT;
            g0 += gj;
            if (g != nullptr) g[j] = -gj;
        }
        acc = block_sum<256>(acc, red);
        g0 = block_sum<256>(g0, red);
        if (threadIdx.x == 0) {
            atomicAdd(loss, acc * norm);
            if (g != nullptr) g[0] = g0;
        }
    }
}

// ------------------------------------------------------------------------------------------ CSR compaction
// Three passes, all parallel over (row, 1024-column segment): count non-zeros per segment, scan the counts (one block),
// ordered fill inside each segment (ballot/popc), so columns stay ascending inside a row (= torch.nonzero order).
constexpr int kSeg = 1024;

__global__ void __launch_bounds__(256)
compact_count_kernel(const float* __restrict__ rep, int V, int first_col, int nseg, int* __restrict__ segcnt) {
    __shared__ float red[8];
    const int b = blockIdx.y, sgm = blockIdx.x;
    const float* row = rep + size_t(b) * V;
    const int c0 = max(first_col, sgm * kSeg), c1 = min(V, (sgm + 1) * kSeg);
    float c = 0.f;
    for (int v = c0 + threadIdx.x; v < c1; v += 256) c += (__ldg(row + v) != 0.f) ? 1.f : 0.f;
    c = block_sum<256>(c, red);
    if (threadIdx.x == 0) segcnt[b * nseg + sgm] = int(c);
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

__global__ void __launch_bounds__(256)
compact_fill_kernel(const float* __restrict__ rep, int V, int first_col, int nseg, const int* __restrict__ offsets,
                    int32_t* __restrict__ cols, float* __restrict__ vals, int capacity,
                    unsigned long long* __restrict__ df_count) {
    __shared__ int warp_cnt[8];
    const int b = blockIdx.y, sgm = blockIdx.x;
    const float* row = rep + size_t(b) * V;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int base = offsets[b * nseg + sgm];
    const int c1 = min(V, (sgm + 1) * kSeg);
    for (int v0 = sgm * kSeg; v0 < c1; v0 += 256) {
        const int v = v0 + threadIdx.x;
        const float x = (v < c1) ? __ldg(row + v) : 0.f;
        // document frequency counts every column; only columns >= first_col are compacted
        if (df_count != nullptr && x > 0.f) atomicAdd(df_count + v, 1ull);
        const bool nz = (v >= first_col) && x != 0.f;
        const uint32_t bal = __ballot_sync(0xffffffffu, nz);
        __syncthreads();
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = base, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (w < warp) off += warp_cnt[w];
            tot += warp_cnt[w];
        }
        off += __popc(bal & ((1u << lane) - 1u));
        if (nz && off < capacity) {
            cols[off] = v;
            vals[off] = x;
        }
        base += tot;
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

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" size_t sb200_scores_workspace_bytes(int Nq, int Nd, int V, int in_batch) {
    (void)Nd; (void)V;
    if (!in_batch || Nq <= 0) return 0;
    return qlists_bytes(Nq);  // thresholded query lists (+ dispatch flag); reused by sb200_scores_bwd
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
    while (tpq > 2 && kRowThreads / tpq < Nq) tpq >>= 1;  // >= 2: lists of up to 64 entries always fit
    return tpq;
}

extern "C" int sb200_scores_fwd(const float* q, const float* d, int Nq, int Nd, int V, int in_batch, float* S,
                                void* workspace, size_t workspace_bytes, sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(q && d && S, "scores_fwd: null pointer");
    SB200_REQUIRE(Nq >= 1 && Nd >= 1 && V >= 1, "scores_fwd: bad shape");
    if (in_batch) {
        const int tj = (Nd + kST - 1) / kST, ti = (Nq + kST - 1) / kST;
        SB200_REQUIRE(ti <= 65535, "scores_fwd: Nq too large");
        const size_t row_smem = row_slab_bytes(V);
        const bool sparse_ok = workspace != nullptr && workspace_bytes >= qlists_bytes(Nq) && row_smem <= 200 * 1024 &&
                               Nq <= 65535;
        const int* dense_flag = nullptr;
        if (sparse_ok) {
            QLists L = qlists_carve(workspace, Nq);
            // flag and the per-row counters are adjacent: one memset
            SB200_CUDA(cudaMemsetAsync(L.flag, 0, 256 + align_up_sz(size_t(Nq) * 4, 256), stream));
            const int tpq = threads_per_query(Nq);
            const int per = kRowThreads / tpq, chunks = (Nq + per - 1) / per;
            const int cap = kEPT * tpq < kQCap ? kEPT * tpq : kQCap;
            q_compact_kernel<<<dim3(8, Nq), 256, 0, stream>>>(q, V, cap, L);
            SB200_CHECK_LAUNCH("q_compact_kernel");
            if (!device_flag_test_and_set(6))
                SB200_CUDA(cudaFuncSetAttribute(scores_docrow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            scores_docrow_kernel<<<dim3(row_kernel_grid(Nd, chunks), chunks), kRowThreads, row_smem, stream>>>(d, Nq, Nd, V, tpq, L, S);
            SB200_CHECK_LAUNCH("scores_docrow_kernel");
            dense_flag = L.flag;
        }
        // dense fp32 tiles: the only path without a workspace, the fallback (device-side flag) with one
        int kchunk;
        const int ks = pick_ksplit(tj * ti, V, kSK, &kchunk);
        if (ks > 1) {
            if (sparse_ok) {
                zero_if_dense_kernel<<<2 * num_sms(), 256, 0, stream>>>(S, size_t(Nq) * Nd, dense_flag);
                SB200_CHECK_LAUNCH("zero_if_dense_kernel");
            } else {
                SB200_CUDA(cudaMemsetAsync(S, 0, size_t(Nq) * Nd * sizeof(float), stream));
            }
        }
        scores_tile_kernel<<<dim3(tj, ti, ks), 256, 0, stream>>>(q, d, Nq, Nd, V, kchunk, dense_flag, S);
        SB200_CHECK_LAUNCH("scores_tile_kernel");
    } else {
        SB200_REQUIRE(Nd % Nq == 0, "scores_fwd: Nd=%d is not a multiple of Nq=%d", Nd, Nq);
        SB200_REQUIRE(Nq <= 65535, "scores_fwd: Nq too large");
        const int G = Nd / Nq;
        int kchunk;
        const int ks = pick_ksplit(Nq, V, 256, &kchunk);
        if (ks > 1) SB200_CUDA(cudaMemsetAsync(S, 0, size_t(Nq) * G * sizeof(float), stream));
        scores_group_kernel<<<dim3(ks, Nq), 256, 0, stream>>>(q, d, G, V, kchunk, S);
        SB200_CHECK_LAUNCH("scores_group_kernel");
    }
    return SB200_OK;
}

extern "C" int sb200_scores_bwd(const float* dS, const float* q, const float* d, int Nq, int Nd, int V, int in_batch,
                                int q_begin, int q_end, int d_begin, int d_end, int accumulate, float* d_q, float* d_d,
                                const void* fwd_workspace, size_t workspace_bytes, sb200_stream_t stream_) {
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
            if (fwd_workspace != nullptr && workspace_bytes >= qlists_bytes(Nq) && row_smem <= 200 * 1024) {
                // query lists built by sb200_scores_fwd: one shared-memory row per document, written to HBM once
                QLists L = qlists_carve(const_cast<void*>(fwd_workspace), Nq);
                if (!device_flag_test_and_set(7))
                    SB200_CUDA(cudaFuncSetAttribute(scores_docrow_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    200 * 1024));
                scores_docrow_bwd_kernel<<<row_kernel_grid(d_end - d_begin, 1), kRowThreads, row_smem, stream>>>(
                    dS, Nq, Nd, V, threads_per_query(Nq), L, d_begin, d_end, accumulate, d_d);
                SB200_CHECK_LAUNCH("scores_docrow_bwd_kernel");
                dense_flag = L.flag;
            }
            dim3 grid(xb, (d_end - d_begin + kBR - 1) / kBR);
            // out row r = doc j, k = query i: coef = dS[i*Nd + j]
            if (vec2) scores_bwd_kernel<2><<<grid, 256, 0, stream>>>(dS, 1, Nd, q, Nq, V, d_begin, d_end, accumulate, dense_flag, d_d);
            else scores_bwd_kernel<1><<<grid, 256, 0, stream>>>(dS, 1, Nd, q, Nq, V, d_begin, d_end, accumulate, dense_flag, d_d);
            SB200_CHECK_LAUNCH("scores_bwd_kernel(d_d)");
        }
        if (d_q != nullptr && q_end > q_begin) {
            dim3 grid(xb, (q_end - q_begin + kBR - 1) / kBR);
            if (vec2) scores_bwd_kernel<2><<<grid, 256, 0, stream>>>(dS, Nd, 1, d, Nd, V, q_begin, q_end, accumulate, nullptr, d_q);
            else scores_bwd_kernel<1><<<grid, 256, 0, stream>>>(dS, Nd, 1, d, Nd, V, q_begin, q_end, accumulate, nullptr, d_q);
            SB200_CHECK_LAUNCH("scores_bwd_kernel(d_q)");
        }
    } else {
        SB200_REQUIRE(Nd % Nq == 0, "scores_bwd: Nd=%d is not a multiple of Nq=%d", Nd, Nq);
        const int G = Nd / Nq;
        int xb = (V + 255) / 256;
        if (xb > 32) xb = 32;
        if (d_d != nullptr && d_end > d_begin) {
            SB200_REQUIRE(d_end - d_begin <= 65535, "scores_bwd: too many rows");
            scores_group_bwd_d_kernel<<<dim3(xb, d_end - d_begin), 256, 0, stream>>>(dS, q, G, V, d_begin, accumulate, d_d);
            SB200_CHECK_LAUNCH("scores_group_bwd_d_kernel");
        }
        if (d_q != nullptr && q_end > q_begin) {
            SB200_REQUIRE(q_end - q_begin <= 65535, "scores_bwd: too many rows");
            scores_group_bwd_q_kernel<<<dim3(xb, q_end - q_begin), 256, 0, stream>>>(dS, d, G, V, q_begin, accumulate, d_q);
            SB200_CHECK_LAUNCH("scores_group_bwd_q_kernel");
        }
    }
    return SB200_OK;
}

extern "C" int sb200_rank_loss(int mode, const float* S, const float* teacher, int Nq, int C, int G, int in_batch,
                               float temperature, float* loss, float* dS, sb200_stream_t stream_) {
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
    SB200_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
    rank_loss_kernel<<<Nq, 256, 0, stream>>>(mode, S, teacher, Nq, C, G, in_batch, 1.f / temperature, loss, dS);
    SB200_CHECK_LAUNCH("rank_loss_kernel");
    return SB200_OK;
}

extern "C" size_t sb200_compact_workspace_bytes(int B, int V) {
    if (B <= 0 || V <= 0) return 0;
    const size_t nseg = size_t((V + kSeg - 1) / kSeg);
    return 2 * align_up(size_t(B) * nseg * sizeof(int), 256);
}

extern "C" int sb200_compact_rows(const float* rep, int B, int V, int first_col, int32_t* row_ptr, int32_t* cols,
                                  float* vals, int capacity, int64_t* df_count, void* workspace, size_t workspace_bytes,
                                  sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(rep && row_ptr && cols && vals, "compact_rows: null pointer");
    SB200_REQUIRE(B >= 1 && B <= 65535 && V >= 1 && first_col >= 0 && capacity >= 0, "compact_rows: bad shape");
    if (workspace == nullptr || workspace_bytes < sb200_compact_workspace_bytes(B, V))
        return fail(SB200_ERR_WORKSPACE, "compact_rows: workspace too small");
    const int nseg = (V + kSeg - 1) / kSeg;
    int* segcnt = static_cast<int*>(workspace);
    int* segoff = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + align_up(size_t(B) * nseg * sizeof(int), 256));
    compact_count_kernel<<<dim3(nseg, B), 256, 0, stream>>>(rep, V, first_col, nseg, segcnt);
    SB200_CHECK_LAUNCH("compact_count_kernel");
    compact_scan_kernel<<<1, 1024, 0, stream>>>(segcnt, B * nseg, nseg, B, segoff, row_ptr);
    SB200_CHECK_LAUNCH("compact_scan_kernel");
    compact_fill_kernel<<<dim3(nseg, B), 256, 0, stream>>>(rep, V, first_col, nseg, segoff, cols, vals, capacity,
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
