// Fused LayerNorm forward / backward for the encoder body (once the sparse head is
// fused, the third-party BERT body dominates the step, and its LayerNorm backward -- PyTorch's gamma/beta reduction
// kernel -- was the largest single item of the r01 launch list: 19.5 % of the C2 step).
//
// Rows = tokens (R = B*L, tens of thousands), H = hidden size (384 / 768 / 1024, H % 128 == 0). HBM-bound:
//   forward : read x, write y                       (2 * R*H*sizeof(T))
//   backward: read x, g, write dx                   (3 * R*H*sizeof(T)) + partial gamma/beta sums (tiny)
// One warp per row, each lane owns H/32 elements as 4-wide vectors (coalesced 8/16-byte accesses), statistics in fp32
// by warp shuffles. The backward kernel accumulates the gamma/beta partial sums of all rows a block walks in registers,
// reduces them across the block's warps in shared memory and writes one partial row per block; a second tiny kernel
// sums the partials in a fixed order (deterministic, no atomics).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.h"

namespace sb200 {
namespace {

constexpr int kLnWarps = 8;           // rows in flight per block
constexpr int kLnThreads = kLnWarps * 32;

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
        const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
        uint2 raw;
        *reinterpret_cast<__nv_bfloat162*>(&raw.x) = __floats2bfloat162_rn(v[0], v[1]);
        *reinterpret_cast<__nv_bfloat162*>(&raw.y) = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = raw;
    }
};

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// NV = H / 128 four-wide vectors per lane; lane's vector k covers columns (lane + 32*k)*4 .. +3
template <typename T, int NV>
__global__ void __launch_bounds__(kLnThreads)
ln_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int R, float eps,
              T* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    constexpr int H = NV * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float gm[NV][4], bt[NV][4];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        Vec4<float>::load(gamma + (lane + 32 * k) * 4, gm[k]);
        Vec4<float>::load(beta + (lane + 32 * k) * 4, bt[k]);
    }
    for (int r = blockIdx.x * kLnWarps + warp; r < R; r += gridDim.x * kLnWarps) {
        const T* xr = x + size_t(r) * H;
        float v[NV][4];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            Vec4<T>::load(xr + (lane + 32 * k) * 4, v[k]);
            s += (v[k][0] + v[k][1]) + (v[k][2] + v[k][3]);
        }
        const float mean = warp_sum(s) * (1.f / H);
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float dlt = v[k][i] - mean;
                q = fmaf(dlt, dlt, q);
            }
        const float rstd = rsqrtf(warp_sum(q) * (1.f / H) + eps);
        T* yr = y + size_t(r) * H;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = fmaf((v[k][i] - mean) * rstd, gm[k][i], bt[k][i]);
            Vec4<T>::store(yr + (lane + 32 * k) * 4, o);
        }
        if (lane == 0) {
            mean_out[r] = mean;
            rstd_out[r] = rstd;
        }
    }
}

template <typename T, int NV>
__global__ void __launch_bounds__(kLnThreads)
ln_bwd_kernel(const T* __restrict__ x, const T* __restrict__ g, const float* __restrict__ gamma,
              const float* __restrict__ mean_in, const float* __restrict__ rstd_in, int R, T* __restrict__ dx,
              float* __restrict__ partial /* [gridDim.x][2][H] */) {
    constexpr int H = NV * 128;
    __shared__ float red[kLnWarps][H];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float gm[NV][4], dg[NV][4], db[NV][4];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        Vec4<float>::load(gamma + (lane + 32 * k) * 4, gm[k]);
#pragma unroll
        for (int i = 0; i < 4; ++i) dg[k][i] = db[k][i] = 0.f;
    }
    for (int r = blockIdx.x * kLnWarps + warp; r < R; r += gridDim.x * kLnWarps) {
        const float mean = __ldg(mean_in + r), rstd = __ldg(rstd_in + r);
        float xh[NV][4], gy[NV][4];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            Vec4<T>::load(x + size_t(r) * H + (lane + 32 * k) * 4, xh[k]);
            Vec4<T>::load(g + size_t(r) * H + (lane + 32 * k) * 4, gy[k]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                xh[k][i] = (xh[k][i] - mean) * rstd;
                dg[k][i] = fmaf(gy[k][i], xh[k][i], dg[k][i]);
                db[k][i] += gy[k][i];
                gy[k][i] *= gm[k][i];          // gradient w.r.t. the normalised value
                s1 += gy[k][i];
                s2 = fmaf(gy[k][i], xh[k][i], s2);
            }
        }
        s1 = warp_sum(s1) * (1.f / H);
        s2 = warp_sum(s2) * (1.f / H);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = rstd * (gy[k][i] - s1 - xh[k][i] * s2);
            Vec4<T>::store(dx + size_t(r) * H + (lane + 32 * k) * 4, o);
        }
    }
    // block-level reduction of the gamma / beta partial sums (two rounds through the same shared buffer)
    float* out = partial + size_t(blockIdx.x) * 2 * H;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NV; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) red[warp][(lane + 32 * k) * 4 + i] = pass == 0 ? dg[k][i] : db[k][i];
        __syncthreads();
        for (int c = threadIdx.x; c < H; c += kLnThreads) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kLnWarps; ++w) s += red[w][c];
            out[pass * H + c] = s;
        }
    }
}

// out[c] = sum_b partial[b][c] over nblocks partial rows of ncols columns, fixed summation order. The work is tiny
// (1-2 MB) and purely latency-bound, so it is spread wide: 8 columns per block, 32 row groups per column (one per lane
// of a warp after the transpose below), 4 independent accumulators per thread, warp-shuffle finish.
__global__ void __launch_bounds__(256)
partial_reduce_kernel(const float* __restrict__ partial, int nblocks, int ncols, float* __restrict__ out0,
                      float* __restrict__ out1, int split) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 8 + warp;  // one warp per column, lanes stride over the partial rows
    if (c >= ncols) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int b = lane;
    for (; b + 96 < nblocks; b += 128) {
        a0 += __ldg(partial + size_t(b) * ncols + c);
        a1 += __ldg(partial + size_t(b + 32) * ncols + c);
        a2 += __ldg(partial + size_t(b + 64) * ncols + c);
        a3 += __ldg(partial + size_t(b + 96) * ncols + c);
    }
    for (; b < nblocks; b += 32) a0 += __ldg(partial + size_t(b) * ncols + c);
    float s = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        if (c < split) out0[c] = s; else out1[c - split] = s;
    }
}

int ln_grid(int R) {
    int g = 4 * num_sms();
    const int need = (R + kLnWarps - 1) / kLnWarps;
    return g < need ? g : need;
}

template <typename T>
int launch_fwd(const void* x, const float* gamma, const float* beta, int R, int H, float eps, void* y, float* mean,
               float* rstd, cudaStream_t stream) {
    const T* xi = static_cast<const T*>(x);
    T* yo = static_cast<T*>(y);
    const int grid = ln_grid(R);
    switch (H / 128) {
        case 1: ln_fwd_kernel<T, 1><<<grid, kLnThreads, 0, stream>>>(xi, gamma, beta, R, eps, yo, mean, rstd); break;
        case 2: ln_fwd_kernel<T, 2><<<grid, kLnThreads, 0, stream>>>(xi, gamma, beta, R, eps, yo, mean, rstd); break;
        case 3: ln_fwd_kernel<T, 3><<<grid, kLnThreads, 0, stream>>>(xi, gamma, beta, R, eps, yo, mean, rstd); break;
        case 4: ln_fwd_kernel<T, 4><<<grid, kLnThreads, 0, stream>>>(xi, gamma, beta, R, eps, yo, mean, rstd); break;
        case 6: ln_fwd_kernel<T, 6><<<grid, kLnThreads, 0, stream>>>(xi, gamma, beta, R, eps, yo, mean, rstd); break;
        case 8: ln_fwd_kernel<T, 8><<<grid, kLnThreads, 0, stream>>>(xi, gamma, beta, R, eps, yo, mean, rstd); break;
        default: return fail(SB200_ERR_ARG, "layer_norm: unsupported H=%d", H);
    }
    SB200_CHECK_LAUNCH("ln_fwd_kernel");
    return SB200_OK;
}

template <typename T>
int launch_bwd(const void* x, const void* g, const float* gamma, const float* mean, const float* rstd, int R, int H,
               void* dx, float* partial, int grid, cudaStream_t stream) {
    const T* xi = static_cast<const T*>(x);
    const T* gi = static_cast<const T*>(g);
    T* dxo = static_cast<T*>(dx);
    switch (H / 128) {
        case 1: ln_bwd_kernel<T, 1><<<grid, kLnThreads, 0, stream>>>(xi, gi, gamma, mean, rstd, R, dxo, partial); break;
        case 2: ln_bwd_kernel<T, 2><<<grid, kLnThreads, 0, stream>>>(xi, gi, gamma, mean, rstd, R, dxo, partial); break;
        case 3: ln_bwd_kernel<T, 3><<<grid, kLnThreads, 0, stream>>>(xi, gi, gamma, mean, rstd, R, dxo, partial); break;
        case 4: ln_bwd_kernel<T, 4><<<grid, kLnThreads, 0, stream>>>(xi, gi, gamma, mean, rstd, R, dxo, partial); break;
        case 6: ln_bwd_kernel<T, 6><<<grid, kLnThreads, 0, stream>>>(xi, gi, gamma, mean, rstd, R, dxo, partial); break;
        case 8: ln_bwd_kernel<T, 8><<<grid, kLnThreads, 0, stream>>>(xi, gi, gamma, mean, rstd, R, dxo, partial); break;
        default: return fail(SB200_ERR_ARG, "layer_norm: unsupported H=%d", H);
    }
    SB200_CHECK_LAUNCH("ln_bwd_kernel");
    return SB200_OK;
}

// ------------------------------------------------------------------------------------------------ residual tail
// out = LayerNorm(dropout(y) + resid) of a transformer block, with the activation delivered twice: fp32 (the residual
// stream autocast keeps in fp32) and bf16 (the operand of the next GEMM). Replaces, per block tail, five PyTorch
// kernels forward (dropout, bf16+fp32 add, LayerNorm, up to three fp32->bf16 casts of the same activation) and the
// matching backward chain (cast-backs, gradient adds, LayerNorm backward, dropout mask multiply). Values are those
// of the stock autocast sequence: dropout result rounded to bf16, add / statistics / affine in fp32, bf16 copy =
// round-to-nearest of the fp32 output. The dropout mask is never stored: Philox4x32-10 keyed by a device-resident
// 64-bit seed and the element index regenerates it in the backward pass.
struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ void operator()(uint32_t c0, uint32_t c1, uint32_t (&out)[4]) const {
        uint32_t c2 = 0u, c3 = 0u, a = k0, b = k1;
#pragma unroll
        for (int round = 0; round < 10; ++round) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ a; c1 = lo1; c2 = hi0 ^ c3 ^ b; c3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

__device__ __forceinline__ float round_bf16(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

struct DropCfg {
    const unsigned long long* seed;  // device pointer, nullptr = no dropout
    uint32_t threshold;              // drop when the 32-bit draw < threshold (= p * 2^32)
    float scale;                     // 1 / (1 - p)
};

// Row layout of the tail kernels: a row of H = NVL * WPR * 128 columns is owned by WPR warps (1 for H <= 512, 2 above:
// with one warp per 768/1024-wide row the backward needs 215-255 registers, one block per SM, 12 % occupancy).
// Warp `half` of the row group holds NVL four-wide vectors per lane: vector k covers columns
// ((half * NVL + k) * 32 + lane) * 4 .. +3. Row statistics of a split row are combined through shared memory and a
// 64-thread named barrier.
template <int NVL, int WPR>
struct RowGeom {
    static constexpr int H = NVL * WPR * 128;
    static constexpr int kRowsPerBlock = kLnWarps / WPR;
    static __device__ __forceinline__ int col(int half, int lane, int k) { return ((half * NVL + k) * 32 + lane) * 4; }
};

// sum of (a, b) over the WPR warps of a row group; xchg = [2][kLnWarps] float2 slots, buf toggles per call
template <int WPR>
__device__ __forceinline__ void row_sum2(float& a, float& b, float2* xchg, int& buf, int warp, int lane) {
    a = warp_sum(a);
    b = warp_sum(b);
    if (WPR == 1) return;
    if (lane == 0) xchg[buf * kLnWarps + warp] = make_float2(a, b);
    asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp >> 1)) : "memory");
    const float2 o = xchg[buf * kLnWarps + (warp ^ 1)];
    buf ^= 1;  // the next exchange uses the other slot set: a slow reader of this one is never overwritten
    a += o.x;
    b += o.y;
}

// branch row r, vector k of this lane: s = dropout(y) + resid (fp32)
template <int NVL, int WPR>
__device__ __forceinline__ void load_branch_sum(const __nv_bfloat16* __restrict__ y, const float* __restrict__ resid, int r,
                                                int half, int lane, bool drop, const Philox& rng, const DropCfg& dc,
                                                float (&s)[NVL][4], uint32_t& keep_bits) {
    using G = RowGeom<NVL, WPR>;
    keep_bits = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < NVL; ++k) {
        const int c = G::col(half, lane, k);
        Vec4<__nv_bfloat16>::load(y + size_t(r) * G::H + c, s[k]);
        if (drop) {
            uint32_t rnd[4];
            const unsigned long long e = (static_cast<unsigned long long>(r) * G::H + c) >> 2;
            rng(static_cast<uint32_t>(e), static_cast<uint32_t>(e >> 32), rnd);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool keep = rnd[i] >= dc.threshold;
                if (!keep) keep_bits &= ~(1u << (k * 4 + i));
                s[k][i] = keep ? round_bf16(s[k][i] * dc.scale) : 0.f;
            }
        }
        if (resid != nullptr) {
            float x[4];
            Vec4<float>::load(resid + size_t(r) * G::H + c, x);
#pragma unroll
            for (int i = 0; i < 4; ++i) s[k][i] += x[i];
        }
    }
}

template <int NVL, int WPR>
__global__ void __launch_bounds__(kLnThreads, NVL <= 3 ? 4 : 3)
add_ln_fwd_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ resid, const float* __restrict__ gamma,
                  const float* __restrict__ beta, int R, float eps, DropCfg dc, float* __restrict__ out32,
                  __nv_bfloat16* __restrict__ out16, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    using G = RowGeom<NVL, WPR>;
    constexpr int H = G::H;
    __shared__ float2 xchg[2 * kLnWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = warp % WPR, slot = warp / WPR;
    int buf = 0;
    float gm[NVL][4], bt[NVL][4];
#pragma unroll
    for (int k = 0; k < NVL; ++k) {
        Vec4<float>::load(gamma + G::col(half, lane, k), gm[k]);
        Vec4<float>::load(beta + G::col(half, lane, k), bt[k]);
    }
    const bool drop = dc.seed != nullptr;
    Philox rng{0u, 0u};
    if (drop) {
        const unsigned long long sd = __ldg(dc.seed);
        rng.k0 = static_cast<uint32_t>(sd);
        rng.k1 = static_cast<uint32_t>(sd >> 32);
    }
    for (int r = blockIdx.x * G::kRowsPerBlock + slot; r < R; r += gridDim.x * G::kRowsPerBlock) {
        float v[NVL][4];
        uint32_t keep_bits;
        load_branch_sum<NVL, WPR>(y, resid, r, half, lane, drop, rng, dc, v, keep_bits);
        float s = 0.f, unused = 0.f;
#pragma unroll
        for (int k = 0; k < NVL; ++k) s += (v[k][0] + v[k][1]) + (v[k][2] + v[k][3]);
        row_sum2<WPR>(s, unused, xchg, buf, warp, lane);
        const float mean = s * (1.f / H);
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < NVL; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float dlt = v[k][i] - mean;
                q = fmaf(dlt, dlt, q);
            }
        unused = 0.f;
        row_sum2<WPR>(q, unused, xchg, buf, warp, lane);
        const float rstd = rsqrtf(q * (1.f / H) + eps);
#pragma unroll
        for (int k = 0; k < NVL; ++k) {
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = fmaf((v[k][i] - mean) * rstd, gm[k][i], bt[k][i]);
            const size_t off = size_t(r) * H + G::col(half, lane, k);
            if (out32 != nullptr) Vec4<float>::store(out32 + off, o);
            if (out16 != nullptr) Vec4<__nv_bfloat16>::store(out16 + off, o);
        }
        if (lane == 0 && half == 0) {
            mean_out[r] = mean;
            rstd_out[r] = rstd;
        }
    }
}

// Gradient of the tail. g32 / g16 are the gradients that arrived at the fp32 and the bf16 copy of the output (either
// may be null); they are summed in fp32. ds (gradient of dropout(y) + resid) leaves as d_resid (fp32) and, through
// the regenerated dropout mask, as d_y (bf16).
template <int NVL, int WPR>
__global__ void __launch_bounds__(kLnThreads, NVL <= 3 ? 3 : 2)
add_ln_bwd_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ resid, const float* __restrict__ g32,
                  const __nv_bfloat16* __restrict__ g16, const float* __restrict__ gamma,
                  const float* __restrict__ mean_in, const float* __restrict__ rstd_in, int R, DropCfg dc,
                  __nv_bfloat16* __restrict__ dy, float* __restrict__ dresid, float* __restrict__ partial) {
    using G = RowGeom<NVL, WPR>;
    constexpr int H = G::H;
    __shared__ float red[G::kRowsPerBlock][H];
    __shared__ float2 xchg[2 * kLnWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = warp % WPR, slot = warp / WPR;
    int buf = 0;
    float gm[NVL][4], dg[NVL][4], db[NVL][4];
#pragma unroll
    for (int k = 0; k < NVL; ++k) {
        Vec4<float>::load(gamma + G::col(half, lane, k), gm[k]);
#pragma unroll
        for (int i = 0; i < 4; ++i) dg[k][i] = db[k][i] = 0.f;
    }
    const bool drop = dc.seed != nullptr;
    Philox rng{0u, 0u};
    if (drop) {
        const unsigned long long sd = __ldg(dc.seed);
        rng.k0 = static_cast<uint32_t>(sd);
        rng.k1 = static_cast<uint32_t>(sd >> 32);
    }
    for (int r = blockIdx.x * G::kRowsPerBlock + slot; r < R; r += gridDim.x * G::kRowsPerBlock) {
        const float mean = __ldg(mean_in + r), rstd = __ldg(rstd_in + r);
        float xh[NVL][4], gy[NVL][4];
        uint32_t keep_bits;
        load_branch_sum<NVL, WPR>(y, resid, r, half, lane, drop, rng, dc, xh, keep_bits);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < NVL; ++k) {
            const size_t off = size_t(r) * H + G::col(half, lane, k);
#pragma unroll
            for (int i = 0; i < 4; ++i) gy[k][i] = 0.f;
            if (g32 != nullptr) Vec4<float>::load(g32 + off, gy[k]);
            if (g16 != nullptr) {
                float t[4];
                Vec4<__nv_bfloat16>::load(g16 + off, t);
#pragma unroll
                for (int i = 0; i < 4; ++i) gy[k][i] += t[i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                xh[k][i] = (xh[k][i] - mean) * rstd;
                dg[k][i] = fmaf(gy[k][i], xh[k][i], dg[k][i]);
                db[k][i] += gy[k][i];
                gy[k][i] *= gm[k][i];
                s1 += gy[k][i];
                s2 = fmaf(gy[k][i], xh[k][i], s2);
            }
        }
        row_sum2<WPR>(s1, s2, xchg, buf, warp, lane);
        s1 *= (1.f / H);
        s2 *= (1.f / H);
#pragma unroll
        for (int k = 0; k < NVL; ++k) {
            const size_t off = size_t(r) * H + G::col(half, lane, k);
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = rstd * (gy[k][i] - s1 - xh[k][i] * s2);
            if (dresid != nullptr) Vec4<float>::store(dresid + off, o);
            if (drop) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    o[i] = ((keep_bits >> (k * 4 + i)) & 1u) ? round_bf16(o[i]) * dc.scale : 0.f;
            }
            Vec4<__nv_bfloat16>::store(dy + off, o);
        }
    }
    float* out = partial + size_t(blockIdx.x) * 2 * H;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NVL; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) red[slot][G::col(half, lane, k) + i] = pass == 0 ? dg[k][i] : db[k][i];
        __syncthreads();
        for (int c = threadIdx.x; c < H; c += kLnThreads) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < G::kRowsPerBlock; ++w) s += red[w][c];
            out[pass * H + c] = s;
        }
    }
}

// one wave of resident blocks (the kernels are latency-bound: every resident warp counts, a second wave only adds a tail)
int tail_grid(int R, int blocks_per_sm, int rows_per_block) {
    const int g = blocks_per_sm * num_sms();
    const int need = (R + rows_per_block - 1) / rows_per_block;
    return g < need ? g : need;
}

DropCfg make_drop(const void* seed, float p) {
    DropCfg dc;
    dc.seed = (seed != nullptr && p > 0.f) ? static_cast<const unsigned long long*>(seed) : nullptr;
    const double t = double(p) * 4294967296.0;
    dc.threshold = t >= 4294967295.0 ? 0xffffffffu : static_cast<uint32_t>(t);
    dc.scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
    return dc;
}

bool ln_supported(int H) {
    const int nv = H / 128;
    return H % 128 == 0 && (nv == 1 || nv == 2 || nv == 3 || nv == 4 || nv == 6 || nv == 8);
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" int sb200_layer_norm_supported(int H) { return ln_supported(H) ? 1 : 0; }

extern "C" int sb200_layer_norm_fwd(const void* x, int elem_bytes, const float* gamma, const float* beta, int R, int H,
                                    float eps, void* y, float* mean, float* rstd, sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(x && gamma && beta && y && mean && rstd, "layer_norm_fwd: null pointer");
    SB200_REQUIRE(R >= 1 && ln_supported(H), "layer_norm_fwd: unsupported shape R=%d H=%d", R, H);
    SB200_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "layer_norm_fwd: elem_bytes=%d (2 = bf16, 4 = fp32)", elem_bytes);
    if (elem_bytes == 2) return launch_fwd<__nv_bfloat16>(x, gamma, beta, R, H, eps, y, mean, rstd, stream);
    return launch_fwd<float>(x, gamma, beta, R, H, eps, y, mean, rstd, stream);
}

extern "C" size_t sb200_layer_norm_bwd_workspace_bytes(int R, int H) {
    if (R <= 0 || H <= 0) return 0;
    // one partial (dgamma, dbeta) row per block of any backward variant: at most 4 blocks per SM, and never more
    // blocks than row groups (>= 4 rows per block: the split-row tail kernels hold kLnWarps / 2 rows per block)
    const int by_rows = (R + kLnWarps / 2 - 1) / (kLnWarps / 2);
    const int cap = 4 * num_sms();
    return size_t(by_rows < cap ? by_rows : cap) * 2 * H * sizeof(float);
}

extern "C" int sb200_layer_norm_bwd(const void* x, const void* dy, int elem_bytes, const float* gamma, const float* mean,
                                    const float* rstd, int R, int H, void* dx, float* dgamma, float* dbeta,
                                    void* workspace, size_t workspace_bytes, sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(x && dy && gamma && mean && rstd && dx && dgamma && dbeta, "layer_norm_bwd: null pointer");
    SB200_REQUIRE(R >= 1 && ln_supported(H), "layer_norm_bwd: unsupported shape R=%d H=%d", R, H);
    SB200_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "layer_norm_bwd: elem_bytes=%d", elem_bytes);
    if (workspace == nullptr || workspace_bytes < sb200_layer_norm_bwd_workspace_bytes(R, H))
        return fail(SB200_ERR_WORKSPACE, "layer_norm_bwd: workspace too small");
    const int grid = ln_grid(R);
    float* partial = static_cast<float*>(workspace);
    int rc = elem_bytes == 2 ? launch_bwd<__nv_bfloat16>(x, dy, gamma, mean, rstd, R, H, dx, partial, grid, stream)
                             : launch_bwd<float>(x, dy, gamma, mean, rstd, R, H, dx, partial, grid, stream);
    if (rc != SB200_OK) return rc;
    partial_reduce_kernel<<<(2 * H + 7) / 8, 256, 0, stream>>>(partial, grid, 2 * H, dgamma, dbeta, H);
    SB200_CHECK_LAUNCH("partial_reduce_kernel");
    return SB200_OK;
}

extern "C" int sb200_add_layer_norm_fwd(const void* y, const float* resid, const float* gamma, const float* beta, int R,
                                        int H, float eps, const void* drop_seed, float drop_p, float* out_f32,
                                        void* out_bf16, float* mean, float* rstd, sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(y && gamma && beta && mean && rstd && (out_f32 || out_bf16), "add_layer_norm_fwd: null pointer");
    SB200_REQUIRE(R >= 1 && ln_supported(H), "add_layer_norm_fwd: unsupported shape R=%d H=%d", R, H);
    SB200_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "add_layer_norm_fwd: drop_p=%f", double(drop_p));
    const DropCfg dc = make_drop(drop_seed, drop_p);
    const __nv_bfloat16* yi = static_cast<const __nv_bfloat16*>(y);
    __nv_bfloat16* o16 = static_cast<__nv_bfloat16*>(out_bf16);
#define SB200_ALN_CASE(NV_, NVL_, WPR_)                                                                               \
    case NV_:                                                                                                         \
        add_ln_fwd_kernel<NVL_, WPR_><<<tail_grid(R, NVL_ <= 3 ? 4 : 3, kLnWarps / WPR_), kLnThreads, 0, stream>>>(   \
            yi, resid, gamma, beta, R, eps, dc, out_f32, o16, mean, rstd);                                            \
        break;
    switch (H / 128) {
        SB200_ALN_CASE(1, 1, 1) SB200_ALN_CASE(2, 2, 1) SB200_ALN_CASE(3, 3, 1) SB200_ALN_CASE(4, 4, 1)
        SB200_ALN_CASE(6, 3, 2) SB200_ALN_CASE(8, 4, 2)
        default: return fail(SB200_ERR_ARG, "add_layer_norm_fwd: unsupported H=%d", H);
    }
#undef SB200_ALN_CASE
    SB200_CHECK_LAUNCH("add_ln_fwd_kernel");
    return SB200_OK;
}

extern "C" int sb200_add_layer_norm_bwd(const void* y, const float* resid, const float* g_f32, const void* g_bf16,
                                        const float* gamma, const float* mean, const float* rstd, int R, int H,
                                        const void* drop_seed, float drop_p, void* d_y, float* d_resid, float* dgamma,
                                        float* dbeta, void* workspace, size_t workspace_bytes, sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(y && gamma && mean && rstd && d_y && dgamma && dbeta && (g_f32 || g_bf16),
                  "add_layer_norm_bwd: null pointer");
    SB200_REQUIRE(R >= 1 && ln_supported(H), "add_layer_norm_bwd: unsupported shape R=%d H=%d", R, H);
    SB200_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "add_layer_norm_bwd: drop_p=%f", double(drop_p));
    if (workspace == nullptr || workspace_bytes < sb200_layer_norm_bwd_workspace_bytes(R, H))
        return fail(SB200_ERR_WORKSPACE, "add_layer_norm_bwd: workspace too small");
    const DropCfg dc = make_drop(drop_seed, drop_p);
    const __nv_bfloat16* yi = static_cast<const __nv_bfloat16*>(y);
    const __nv_bfloat16* g16 = static_cast<const __nv_bfloat16*>(g_bf16);
    __nv_bfloat16* dyo = static_cast<__nv_bfloat16*>(d_y);
    float* partial = static_cast<float*>(workspace);
    int grid = 0;  // <= ln_grid(R) blocks: the workspace holds one partial row per block
#define SB200_ALN_CASE(NV_, NVL_, WPR_)                                                                               \
    case NV_:                                                                                                         \
        grid = tail_grid(R, NVL_ <= 3 ? 3 : 2, kLnWarps / WPR_);                                                      \
        if (size_t(grid) * 2 * H * sizeof(float) > workspace_bytes)                                                   \
            return fail(SB200_ERR_WORKSPACE, "add_layer_norm_bwd: workspace too small for %d blocks", grid);          \
        add_ln_bwd_kernel<NVL_, WPR_><<<grid, kLnThreads, 0, stream>>>(yi, resid, g_f32, g16, gamma, mean, rstd, R,   \
                                                                       dc, dyo, d_resid, partial);                    \
        break;
    switch (H / 128) {
        SB200_ALN_CASE(1, 1, 1) SB200_ALN_CASE(2, 2, 1) SB200_ALN_CASE(3, 3, 1) SB200_ALN_CASE(4, 4, 1)
        SB200_ALN_CASE(6, 3, 2) SB200_ALN_CASE(8, 4, 2)
        default: return fail(SB200_ERR_ARG, "add_layer_norm_bwd: unsupported H=%d", H);
    }
#undef SB200_ALN_CASE
    SB200_CHECK_LAUNCH("add_ln_bwd_kernel");
    partial_reduce_kernel<<<(2 * H + 7) / 8, 256, 0, stream>>>(partial, grid, 2 * H, dgamma, dbeta, H);
    SB200_CHECK_LAUNCH("partial_reduce_kernel");
    return SB200_OK;
}

// ------------------------------------------------------------------------------------------------ bias gradient
// db[c] = sum_r dy[r, c] for a row-major [R, N] gradient (bf16 or fp32): the bias gradient of every Linear of the body.
// PyTorch's generic reduce kernel took 58 us per call on [40960, 384..1536] bf16 (11.6 % of the C2 step after the
// LayerNorm fusion); this is one streaming pass with 16-byte loads, per-block partial rows and a fixed-order finish.
namespace sb200 {
namespace {

constexpr int kCsWarps = 8;
constexpr int kCsMaxVec = 16;  // N <= 16 * 256 = 4096 columns

template <typename T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
};
template <> struct Vec8<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};

template <typename T, int NV>
__global__ void __launch_bounds__(kCsWarps * 32)
colsum_partial_kernel(const T* __restrict__ dy, int R, int N, float* __restrict__ partial /* [gridDim.x][N] */) {
    extern __shared__ float red[];  // [kCsWarps][N]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvec = N >> 3;
    float acc[NV][8];
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
    const int stride = gridDim.x * kCsWarps;
    constexpr int U = NV <= 2 ? 4 : (NV <= 4 ? 2 : 1);  // rows in flight per warp (register budget)
    for (int r = blockIdx.x * kCsWarps + warp; r < R; r += U * stride) {
        // U independent rows per iteration keep U*NV 16-byte loads in flight per lane
        float v[U][NV][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int rr = r + u * stride;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const int vc = lane + 32 * k;
                if (rr < R && vc < nvec) {
                    Vec8<T>::load(dy + size_t(rr) * N + vc * 8, v[u][k]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[u][k][i] = 0.f;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < NV; ++k)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[k][i] += v[u][k][i];
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int vc = lane + 32 * k;
        if (vc < nvec)
#pragma unroll
            for (int i = 0; i < 8; ++i) red[warp * N + vc * 8 + i] = acc[k][i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < N; c += kCsWarps * 32) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kCsWarps; ++w) s += red[w * N + c];
        partial[size_t(blockIdx.x) * N + c] = s;
    }
}

// 16-byte vectors per lane: the smallest instantiated count that covers N columns (lanes past N/8 are masked)
int colsum_nv(int N) {
    const int need = (N / 8 + 31) / 32;
    const int have[] = {1, 2, 3, 4, 6, 8, 12, 16};
    for (int v : have)
        if (v >= need) return v;
    return 0;
}

int colsum_grid(int R) {
    int g = 2 * num_sms();
    const int need = (R + kCsWarps - 1) / kCsWarps;
    return g < need ? g : need;
}

template <typename T>
int launch_colsum(const void* dy, int R, int N, float* partial, int grid, cudaStream_t stream) {
    const T* p = static_cast<const T*>(dy);
    const size_t smem = size_t(kCsWarps) * N * sizeof(float);
    const int nv = colsum_nv(N);
#define SB200_CS_CASE(NV_)                                                                                          \
    case NV_:                                                                                                       \
        if (smem > 48 * 1024 && !device_flag_test_and_set(8 + (NV_ == 16 ? 2 : (NV_ == 12 ? 1 : 0)) + (sizeof(T) == 4 ? 3 : 0)))                                                       \
            SB200_CUDA(cudaFuncSetAttribute(colsum_partial_kernel<T, NV_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                            int(kCsWarps * 4096 * sizeof(float))));                                 \
        colsum_partial_kernel<T, NV_><<<grid, kCsWarps * 32, smem, stream>>>(p, R, N, partial);                     \
        break;
    switch (nv) {
        SB200_CS_CASE(1) SB200_CS_CASE(2) SB200_CS_CASE(3) SB200_CS_CASE(4) SB200_CS_CASE(6) SB200_CS_CASE(8)
        SB200_CS_CASE(12) SB200_CS_CASE(16)
        default: return fail(SB200_ERR_ARG, "colsum: unsupported N=%d", N);
    }
#undef SB200_CS_CASE
    SB200_CHECK_LAUNCH("colsum_partial_kernel");
    return SB200_OK;
}

bool colsum_supported(int N) {
    return N >= 8 && N % 8 == 0 && N <= 4096;
}

}  // namespace
}  // namespace sb200

extern "C" int sb200_colsum_supported(int N) { return sb200::colsum_supported(N) ? 1 : 0; }

extern "C" size_t sb200_colsum_workspace_bytes(int R, int N) {
    if (R <= 0 || N <= 0) return 0;
    return size_t(sb200::colsum_grid(R)) * N * sizeof(float);
}

extern "C" int sb200_colsum(const void* dy, int elem_bytes, int R, int N, float* out, void* workspace,
                            size_t workspace_bytes, sb200_stream_t stream_) {
    using namespace sb200;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(dy && out, "colsum: null pointer");
    SB200_REQUIRE(R >= 1 && colsum_supported(N), "colsum: unsupported shape R=%d N=%d", R, N);
    SB200_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "colsum: elem_bytes=%d", elem_bytes);
    if (workspace == nullptr || workspace_bytes < sb200_colsum_workspace_bytes(R, N))
        return fail(SB200_ERR_WORKSPACE, "colsum: workspace too small");
    const int grid = colsum_grid(R);
    float* partial = static_cast<float*>(workspace);
    const int rc = elem_bytes == 2 ? launch_colsum<__nv_bfloat16>(dy, R, N, partial, grid, stream)
                                   : launch_colsum<float>(dy, R, N, partial, grid, stream);
    if (rc != SB200_OK) return rc;
    partial_reduce_kernel<<<(N + 7) / 8, 256, 0, stream>>>(partial, grid, N, out, out, N);
    SB200_CHECK_LAUNCH("partial_reduce_kernel");
    return SB200_OK;
}

// ------------------------------------------------------------------------------------------------ GELU
// y = x * Phi(x) (the exact "gelu" of transformers' BertIntermediate / BertPredictionHeadTransform) on bf16 rows, and
// its backward fused with the bias gradient of the Linear in front of it: dx = dy * (Phi(x) + x * phi(x)),
// db[c] = sum_r dx[r, c]. Phi through erfc by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below bf16
// resolution; one ex2 and one rcp per element instead of erff's long polynomial, which keeps both kernels
// HBM-bound: forward 4 B/element, backward 6 B/element).
namespace sb200 {
namespace {

// ~15 issue slots per element (the first version, with __expf / __frcp_rn and their special-case paths, needed 37
// and made both kernels issue-bound at the speed of PyTorch's erff kernels: 53.5 M elements per call).
__device__ __forceinline__ float ex2_approx(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float rcp_approx(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ void gelu_terms(float x, float& cdf, float& pdf_x) {
    const float e = ex2_approx(x * x * -0.72134752044448170f);               // exp(-x^2 / 2)
    const float t = rcp_approx(fmaf(0.23164189798f, fabsf(x), 1.f));         // 1 / (1 + 0.3275911 * |x| / sqrt 2)
    float p = fmaf(0.5307027145f, t, -0.7265760135f);                        // 0.5 * A&S coefficients
    p = fmaf(p, t, 0.7107068705f);
    p = fmaf(p, t, -0.142248368f);
    p = fmaf(p, t, 0.127414796f);
    const float half_erfc = p * t * e;                                       // 0.5 * erfc(|x| / sqrt 2)
    cdf = 0.5f + copysignf(0.5f - half_erfc, x);                             // Phi(x) = 1 - Phi(-x)
    pdf_x = x * e * 0.39894228040143268f;                                    // x * phi(x)
}

__device__ __forceinline__ void unpack8(const uint4& raw, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    uint4 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    return raw;
}

constexpr int kGeluThreads = 256;
constexpr int kGeluUnroll = 4;  // 16-byte vectors in flight per thread

__global__ void __launch_bounds__(kGeluThreads)
gelu_fwd_kernel(const uint4* __restrict__ x, size_t nvec, uint4* __restrict__ y) {
    const size_t stride = size_t(gridDim.x) * kGeluThreads;
    for (size_t base = size_t(blockIdx.x) * kGeluThreads + threadIdx.x; base < nvec; base += stride * kGeluUnroll) {
        uint4 raw[kGeluUnroll];
#pragma unroll
        for (int u = 0; u < kGeluUnroll; ++u)
            if (base + u * stride < nvec) raw[u] = __ldg(x + base + u * stride);
#pragma unroll
        for (int u = 0; u < kGeluUnroll; ++u) {
            if (base + u * stride >= nvec) continue;
            float v[8];
            unpack8(raw[u], v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float cdf, px;
                gelu_terms(v[i], cdf, px);
                v[i] *= cdf;
            }
            y[base + u * stride] = pack8(v);
        }
    }
}

// Each warp owns one 256-column slab (blockIdx.y; a lane's 16-byte vector = 8 columns) and walks rows with kGbRows
// of them in flight: few registers (8 bias-gradient accumulators per lane), all blocks resident, 2*kGbRows 16-byte
// loads outstanding per lane. (A first version with one warp per full row kept N/32 accumulators per lane: 104
// registers, 25 % occupancy, 3.4 TB/s.)
constexpr int kGbRows = 4;

template <bool kColsum>
__global__ void __launch_bounds__(kCsWarps * 32, 4)
gelu_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, int R, int N,
                __nv_bfloat16* __restrict__ dx, float* __restrict__ partial /* [gridDim.x][N] */) {
    __shared__ float red[kCsWarps][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = blockIdx.y * 256 + lane * 8;
    const bool live = col < N;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const int stride = gridDim.x * kCsWarps;
    for (int r = blockIdx.x * kCsWarps + warp; r < R; r += kGbRows * stride) {
        uint4 xr[kGbRows], gr[kGbRows];
#pragma unroll
        for (int u = 0; u < kGbRows; ++u) {
            const int rr = r + u * stride;
            const bool ok = live && rr < R;
            xr[u] = ok ? __ldg(reinterpret_cast<const uint4*>(x + size_t(rr) * N + col)) : make_uint4(0u, 0u, 0u, 0u);
            gr[u] = ok ? __ldg(reinterpret_cast<const uint4*>(dy + size_t(rr) * N + col)) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < kGbRows; ++u) {
            const int rr = r + u * stride;
            float xv[8], gv[8];
            unpack8(xr[u], xv);
            unpack8(gr[u], gv);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float cdf, px;
                gelu_terms(xv[i], cdf, px);
                gv[i] *= cdf + px;
            }
            const uint4 out = pack8(gv);
            if (live && rr < R) *reinterpret_cast<uint4*>(dx + size_t(rr) * N + col) = out;
            if (kColsum) {
                // the bias gradient sums the bf16 values the GEMMs will read, like the unfused path
                // (rows past R contribute gelu'(0) * 0 = 0)
                unpack8(out, gv);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += gv[i];
            }
        }
    }
    if (!kColsum) return;
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
    __syncthreads();
    const int c = threadIdx.x;  // 256 threads = 256 columns of the slab
    if (blockIdx.y * 256 + c < N) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kCsWarps; ++w) s += red[w][c];
        partial[size_t(blockIdx.x) * N + blockIdx.y * 256 + c] = s;
    }
}

int gelu_bwd_row_blocks(int R, int N) {
    const int nslab = (N + 255) / 256;
    int gx = (4 * num_sms() + nslab - 1) / nslab;          // ~4 resident blocks per SM in total
    const int need = (R + kCsWarps - 1) / kCsWarps;
    if (gx > need) gx = need;
    const int cap = colsum_grid(R);                          // the workspace is sized for this many partial rows
    return gx < cap ? (gx < 1 ? 1 : gx) : cap;
}

}  // namespace
}  // namespace sb200

extern "C" int sb200_gelu_fwd(const void* x, size_t n, void* y, sb200_stream_t stream_) {
    using namespace sb200;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(x && y, "gelu_fwd: null pointer");
    SB200_REQUIRE(n >= 8 && n % 8 == 0, "gelu_fwd: n=%zu must be a positive multiple of 8", n);
    const size_t nvec = n / 8;
    const size_t per_block = size_t(kGeluThreads) * kGeluUnroll;
    size_t grid = (nvec + per_block - 1) / per_block;
    const size_t cap = size_t(num_sms()) * 8;
    if (grid > cap) grid = cap;
    gelu_fwd_kernel<<<int(grid), kGeluThreads, 0, stream>>>(static_cast<const uint4*>(x), nvec, static_cast<uint4*>(y));
    SB200_CHECK_LAUNCH("gelu_fwd_kernel");
    return SB200_OK;
}

extern "C" size_t sb200_gelu_bwd_workspace_bytes(int R, int N) { return sb200_colsum_workspace_bytes(R, N); }

extern "C" int sb200_gelu_bwd(const void* x, const void* dy, int R, int N, void* dx, float* colsum, void* workspace,
                              size_t workspace_bytes, sb200_stream_t stream_) {
    using namespace sb200;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(x && dy && dx, "gelu_bwd: null pointer");
    SB200_REQUIRE(R >= 1 && colsum_supported(N), "gelu_bwd: unsupported shape R=%d N=%d", R, N);
    float* partial = nullptr;
    if (colsum != nullptr) {
        if (workspace == nullptr || workspace_bytes < sb200_gelu_bwd_workspace_bytes(R, N))
            return fail(SB200_ERR_WORKSPACE, "gelu_bwd: workspace too small");
        partial = static_cast<float*>(workspace);
    }
    const int gx = gelu_bwd_row_blocks(R, N);
    const dim3 grid3(gx, (N + 255) / 256);
    const int grid = gx;
    const __nv_bfloat16* xi = static_cast<const __nv_bfloat16*>(x);
    const __nv_bfloat16* gi = static_cast<const __nv_bfloat16*>(dy);
    __nv_bfloat16* dxo = static_cast<__nv_bfloat16*>(dx);
    if (partial != nullptr)
        gelu_bwd_kernel<true><<<grid3, kCsWarps * 32, 0, stream>>>(xi, gi, R, N, dxo, partial);
    else
        gelu_bwd_kernel<false><<<grid3, kCsWarps * 32, 0, stream>>>(xi, gi, R, N, dxo, nullptr);
    SB200_CHECK_LAUNCH("gelu_bwd_kernel");
    if (partial != nullptr) {
        partial_reduce_kernel<<<(N + 7) / 8, 256, 0, stream>>>(partial, grid, N, colsum, colsum, N);
        SB200_CHECK_LAUNCH("partial_reduce_kernel");
    }
    return SB200_OK;
}
