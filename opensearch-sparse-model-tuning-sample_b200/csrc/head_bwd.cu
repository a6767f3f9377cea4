// Sparse head backward for sm_100a.
//
// Reference: autograd backward of scripts/model/sparse_encoders.py:108-114. There it materialises a dense
// dlogits[B,L,V] (non-zero only at the B*V arg-max positions) and runs two dense GEMMs. Here the B*V coefficients
//     c[b,v] = d_rep[b,v] * f'(xmax[b,v]),   f = log1p o relu  (o log1p with use_l0)
// are applied directly at the winning positions l* = argmax[b,v]:
//     dW[v,:]        = sum_b c[b,v] * hidden[b, l*(b,v), :]       (gather, one owner per v -> no atomics)
//     dbias[v]       = sum_b c[b,v]
//     d_hidden[b,l,:] = sum_{v: l*(b,v)=l} c[b,v] * W[v,:]        (entries bucketed by l, then gathered)
// B*V*H multiply-adds instead of 2*B*L*V*H, and entries with c == 0 (inactive vocabulary) cost nothing.
// All three kernels are gather-bound on L2 (hidden and W are L2-resident: tens of MB).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.h"

namespace sb200 {
namespace {

constexpr int kMaxChunks = 4;          // H <= 1024: 16-byte chunks per lane
constexpr int kDwRows = 32;            // vocab rows per dW block
constexpr int kDwThreads = 256;
constexpr int kDhEntriesPerBlock = 256;
constexpr int kBucketThreads = 1024;
constexpr int kMaxL = 4096;

__device__ __forceinline__ float head_coef(float g, float x, int l0) {
    // matches torch autograd order: grad/(1+r1) [l0], then /(1+relu(x)), then relu mask (0 at x <= 0)
    if (!(x > 0.f) || g == 0.f) return 0.f;
    float c = g;
    if (l0) c = c / (1.f + log1pf(x));
    return c / (1.f + x);
}

// acc[0..8) += c * (8 half-precision values in `raw`): bf16 or, kFp16, IEEE fp16
template <bool kFp16>
__device__ __forceinline__ void fma_half8(float (&acc)[8], const uint4& raw, float c) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f;
        if (kFp16) f = __half22float2(reinterpret_cast<const __half2*>(&raw)[i]);
        else f = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(&raw)[i]);
        acc[2 * i] = fmaf(c, f.x, acc[2 * i]);
        acc[2 * i + 1] = fmaf(c, f.y, acc[2 * i + 1]);
    }
}

// ---------------------------------------------------------------- dW / dbias
// Block = 32 consecutive vocab rows; (c, l*) for [Bc sequences x 32 rows] staged in smem (coalesced),
// then each warp owns 4 rows and walks the sequences, gathering hidden rows with 16-byte loads.
template <int NCHUNK, bool kFp16>
__global__ void __launch_bounds__(kDwThreads)
bwd_dw_kernel(const float* __restrict__ d_rep, const float* __restrict__ xmax, const int32_t* __restrict__ argmax,
              const __nv_bfloat16* __restrict__ hidden, int B, int L, int H, int V, int l0, int Bc,
              float* __restrict__ dW, float* __restrict__ dbias) {
    extern __shared__ float2 stage[];  // [Bc][32] (c, l as int bits)
    const int v0 = blockIdx.x * kDwRows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks = H >> 3;

    for (int b0 = 0; b0 < B; b0 += Bc) {
        const int nb = min(Bc, B - b0);
        __syncthreads();
        for (int t = threadIdx.x; t < nb * kDwRows; t += kDwThreads) {
            const int bo = t >> 5, vo = t & 31;
            const int v = v0 + vo;
            float c = 0.f;
            int l = 0;
            if (v < V) {
                const size_t o = size_t(b0 + bo) * V + v;
                c = head_coef(__ldg(d_rep + o), __ldg(xmax + o), l0);
                l = __ldg(argmax + o);
                l = min(max(l, 0), L - 1);
            }
            stage[t] = make_float2(c, __int_as_float(l));
        }
        __syncthreads();
        for (int r = 0; r < 4; ++r) {
            const int vo = warp * 4 + r;
            const int v = v0 + vo;
            if (v >= V) break;  // warp-uniform
            float acc[NCHUNK][8];
            float bsum = 0.f;
#pragma unroll
            for (int k = 0; k < NCHUNK; ++k)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
#pragma unroll 4
            for (int bo = 0; bo < nb; ++bo) {
                const float2 e = stage[bo * kDwRows + vo];  // broadcast
                if (e.x == 0.f) continue;                   // warp-uniform
                bsum += e.x;
                const __nv_bfloat16* row = hidden + (size_t(b0 + bo) * L + __float_as_int(e.y)) * H;
#pragma unroll
                for (int k = 0; k < NCHUNK; ++k) {
                    const int ch = lane + 32 * k;
                    if (ch < n_chunks) {
                        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(row) + ch);
                        fma_half8<kFp16>(acc[k], raw, e.x);
                    }
                }
            }
            // this block is the only writer of rows v0..v0+31: first pass stores, later passes accumulate
#pragma unroll
            for (int k = 0; k < NCHUNK; ++k) {
                const int ch = lane + 32 * k;
                if (ch < n_chunks) {
                    float4* dst = reinterpret_cast<float4*>(dW + size_t(v) * H + ch * 8);
                    float4 lo = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
                    float4 hi = make_float4(acc[k][4], acc[k][5], acc[k][6], acc[k][7]);
                    if (b0 > 0) {
                        const float4 plo = dst[0], phi = dst[1];
                        lo.x += plo.x; lo.y += plo.y; lo.z += plo.z; lo.w += plo.w;
                        hi.x += phi.x; hi.y += phi.y; hi.z += phi.z; hi.w += phi.w;
                    }
                    dst[0] = lo;
                    dst[1] = hi;
                }
            }
            if (dbias != nullptr && lane == 0) dbias[v] = (b0 > 0 ? dbias[v] : 0.f) + bsum;
        }
    }
}

// ---------------------------------------------------------------- bucket active entries of one sequence by l*
// entries[b][pos] = (key = l << 20 | v, c), sorted by l (order inside a bucket is arbitrary); nact[b] = count.
__global__ void __launch_bounds__(kBucketThreads)
bwd_bucket_kernel(const float* __restrict__ d_rep, const float* __restrict__ xmax, const int32_t* __restrict__ argmax,
                  int L, int V, int l0, uint2* __restrict__ entries, int* __restrict__ nact) {
    __shared__ int cnt[kMaxL];
    __shared__ int warp_tot[32];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int l = tid; l < L; l += kBucketThreads) cnt[l] = 0;
    __syncthreads();
    const size_t base = size_t(b) * V;
    // pass 1: entries per position. Only "is the coefficient non-zero" matters here (x > 0 and g != 0, see head_coef);
    // lanes of a warp that hit the same counter are aggregated into one shared-memory atomic (dense regime: ~120
    // entries per counter).
    for (int v0 = 0; v0 < V; v0 += kBucketThreads) {
        const int v = v0 + tid;
        bool active = false;
        int l = 0;
        if (v < V) {
            active = (__ldg(xmax + base + v) > 0.f) && (__ldg(d_rep + base + v) != 0.f);
            if (active) l = min(max(__ldg(argmax + base + v), 0), L - 1);
        }
        const unsigned int peers = __match_any_sync(0xffffffffu, active ? l : -1);
        if (active && lane == __ffs(peers) - 1) atomicAdd(&cnt[l], __popc(peers));
    }
    __syncthreads();
    // exclusive scan of cnt[0..L) in place; each thread owns a contiguous run of 4 counters
    {
        constexpr int per = kMaxL / kBucketThreads;  // 4
        int vals[per];
        int sum = 0;
#pragma unroll
        for (int i = 0; i < per; ++i) {
            const int l = tid * per + i;
            vals[i] = (l < L) ? cnt[l] : 0;
            sum += vals[i];
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_tot[lane] = wi - w;  // exclusive
            if (lane == 31) nact[b] = wi;
        }
        __syncthreads();
        int run = warp_tot[warp] + incl - sum;
#pragma unroll
        for (int i = 0; i < per; ++i) {
            const int l = tid * per + i;
            if (l < L) cnt[l] = run;
            run += vals[i];
        }
    }
    __syncthreads();
    uint2* out = entries + base;
    for (int v0 = 0; v0 < V; v0 += kBucketThreads) {
        const int v = v0 + tid;
        float c = 0.f;
        int l = 0;
        bool active = false;
        if (v < V) {
            const float gv = __ldg(d_rep + base + v), xv = __ldg(xmax + base + v);
            active = (xv > 0.f) && (gv != 0.f);        // the same predicate as pass 1 (a coefficient that underflows to
            if (active) {                              // zero keeps its slot and contributes nothing)
                c = head_coef(gv, xv, l0);
                l = min(max(__ldg(argmax + base + v), 0), L - 1);
            }
        }
        const unsigned int peers = __match_any_sync(0xffffffffu, active ? l : -1);
        const int leader = __ffs(peers) - 1;
        int slot = 0;
        if (active && lane == leader) slot = atomicAdd(&cnt[l], __popc(peers));
        slot = __shfl_sync(0xffffffffu, slot, leader);
        if (active) out[slot + __popc(peers & ((1u << lane) - 1u))] = make_uint2((uint32_t(l) << 20) | uint32_t(v), __float_as_uint(c));
    }
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---------------------------------------------------------------- d_hidden
// grid (ceil(V / 256), B). Each warp takes 32 consecutive bucketed entries, accumulates runs of equal l in
// registers and flushes a run with vector reductions into the (pre-zeroed) fp32 d_hidden row.
template <int NCHUNK, bool kFp16>
__global__ void __launch_bounds__(256)
bwd_dh_kernel(const uint2* __restrict__ entries, const int* __restrict__ nact, const __nv_bfloat16* __restrict__ W,
              int L, int H, int V, float* __restrict__ d_hidden) {
    const int b = blockIdx.y;
    const int n = __ldg(nact + b);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int first = blockIdx.x * kDhEntriesPerBlock + warp * 32;
    if (first >= n) return;
    const int cnt = min(32, n - first);
    const int n_chunks = H >> 3;
    uint2 mine = make_uint2(0u, 0u);
    if (lane < cnt) mine = __ldg(entries + size_t(b) * V + first + lane);

    float acc[NCHUNK][8];
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;

    auto flush = [&](int l) {
        float* row = d_hidden + (size_t(b) * L + l) * H;
#pragma unroll
        for (int k = 0; k < NCHUNK; ++k) {
            const int ch = lane + 32 * k;
            if (ch < n_chunks) {
                red_add_v4(row + ch * 8, acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
                red_add_v4(row + ch * 8 + 4, acc[k][4], acc[k][5], acc[k][6], acc[k][7]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
        }
    };

    int cur_l = int(__shfl_sync(0xffffffffu, mine.x, 0) >> 20);
    for (int e0 = 0; e0 < cnt; e0 += 4) {
        uint4 raw[4][NCHUNK];
        uint32_t key[4];
        float cf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = min(e0 + j, cnt - 1);
            key[j] = __shfl_sync(0xffffffffu, mine.x, e);
            cf[j] = (e0 + j < cnt) ? __uint_as_float(__shfl_sync(0xffffffffu, mine.y, e)) : 0.f;
            const __nv_bfloat16* row = W + size_t(key[j] & 0xFFFFFu) * H;
#pragma unroll
            for (int k = 0; k < NCHUNK; ++k) {
                const int ch = lane + 32 * k;
                raw[j][k] = (ch < n_chunks) ? __ldg(reinterpret_cast<const uint4*>(row) + ch) : make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (e0 + j < cnt) {
                const int l = int(key[j] >> 20);
                if (l != cur_l) {
                    flush(cur_l);
                    cur_l = l;
                }
#pragma unroll
                for (int k = 0; k < NCHUNK; ++k) fma_half8<kFp16>(acc[k], raw[j][k], cf[j]);
            }
        }
    }
    flush(cur_l);
}

// ---------------------------------------------------------------- exact-fit variants (H * 2 bytes = VB * 32 * NCH)
// The generic kernels above predicate every 16-byte chunk on `ch < n_chunks` (H = 384 leaves half of the warp idle on
// the second chunk) and carry 64-bit address arithmetic and the run-flush control flow through every entry: ncu showed
// them issue-bound (94 / 66 warp instructions per entry, IPC 2.1-2.5, L2 at 21-31 %). For the common widths each lane
// instead owns NCH chunks of VB bytes that tile the row exactly, the run structure of a warp's 32 entries comes from
// one ballot, and rows are fetched four entries ahead of their use.
template <int VB> struct Chunk;
template <> struct Chunk<16> { using type = uint4; static constexpr int kWords = 4; };
template <> struct Chunk<8> { using type = uint2; static constexpr int kWords = 2; };
template <> struct Chunk<4> { using type = uint32_t; static constexpr int kWords = 1; };

template <bool kFp16>
__device__ __forceinline__ void fma_word(float& a0, float& a1, uint32_t w, float c) {
    float lo, hi;
    if (kFp16) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
        lo = f.x;
        hi = f.y;
    } else {
        lo = __uint_as_float(w << 16);
        hi = __uint_as_float(w & 0xffff0000u);
    }
    a0 = fmaf(c, lo, a0);
    a1 = fmaf(c, hi, a1);
}
template <int VB, bool kFp16>
__device__ __forceinline__ void fma_chunk(float* acc, const typename Chunk<VB>::type& raw, float c) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&raw);
#pragma unroll
    for (int i = 0; i < Chunk<VB>::kWords; ++i) fma_word<kFp16>(acc[2 * i], acc[2 * i + 1], w[i], c);
}
// the lane's chunks of one row: chunk k lives at byte offset (lane + 32 k) * VB
template <int VB, int NCH>
__device__ __forceinline__ void load_row(typename Chunk<VB>::type (&dst)[NCH], const char* row_lane) {
#pragma unroll
    for (int k = 0; k < NCH; ++k)
        dst[k] = __ldg(reinterpret_cast<const typename Chunk<VB>::type*>(row_lane + k * 32 * VB));
}

template <int VB, int NCH, bool kFp16>
__global__ void __launch_bounds__(kDwThreads)
bwd_dw_fit_kernel(const float* __restrict__ d_rep, const float* __restrict__ xmax, const int32_t* __restrict__ argmax,
                  const char* __restrict__ hidden, const int32_t* __restrict__ cu, int B, int L, int V, int l0, int Bc,
                  float* __restrict__ dW, float* __restrict__ dbias) {
    constexpr int kRowBytes = VB * 32 * NCH;       // = H * 2
    constexpr int kPer = VB / 2;                    // fp32 accumulators per chunk
    extern __shared__ float2 stage[];               // [32 vocab rows][Bc + 1] (c, row index as int bits)
    const int v0 = blockIdx.x * kDwRows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pitch = Bc + 1;                       // odd pitch: the transposed fill is at most 2-way bank conflicted
    for (int b0 = 0; b0 < B; b0 += Bc) {
        const int nb = min(Bc, B - b0);
        __syncthreads();
        for (int t = threadIdx.x; t < nb * kDwRows; t += kDwThreads) {
            const int bo = t >> 5, vo = t & 31;
            const int v = v0 + vo;
            float c = 0.f;
            int l = 0;
            if (v < V) {
                const size_t o = size_t(b0 + bo) * V + v;
                c = head_coef(__ldg(d_rep + o), __ldg(xmax + o), l0);
                l = min(max(__ldg(argmax + o), 0), L - 1);
            }
            // absolute row of the gathered hidden state: padded [B, L, H] or packed [T, H] (cu = sequence starts)
            const int row = (cu != nullptr) ? min(__ldg(cu + b0 + bo) + l, __ldg(cu + b0 + bo + 1) - 1) : (b0 + bo) * L + l;
            stage[vo * pitch + bo] = make_float2(c, __int_as_float(max(row, 0)));
        }
        __syncthreads();
        const char* base = hidden + lane * VB;
        for (int r = 0; r < 4; ++r) {
            const int vo = warp * 4 + r;
            const int v = v0 + vo;
            if (v >= V) break;  // warp-uniform
            float acc[NCH][kPer];
            float bsum = 0.f;
#pragma unroll
            for (int k = 0; k < NCH; ++k)
#pragma unroll
                for (int i = 0; i < kPer; ++i) acc[k][i] = 0.f;
            for (int bo0 = 0; bo0 < nb; bo0 += 32) {
                // 32 sequences at a time: lane i holds sequence bo0+i's (c, row); entries with c == 0 (inactive vocabulary
                // in that sequence -- almost all of them once the model is trained) are skipped through the ballot
                float2 mine = make_float2(0.f, 0.f);
                if (bo0 + lane < nb) mine = stage[vo * pitch + bo0 + lane];
                bsum += mine.x;
                uint32_t live = __ballot_sync(0xffffffffu, mine.x != 0.f);
                while (live != 0u) {
                    int src[4];
                    float cf[4];
                    typename Chunk<VB>::type raw[4][NCH];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {       // up to four gathered rows in flight
                        src[u] = live != 0u ? (__ffs(live) - 1) : -1;
                        if (live != 0u) live &= live - 1;
                        const int sl = src[u] < 0 ? 0 : src[u];
                        cf[u] = src[u] < 0 ? 0.f : __shfl_sync(0xffffffffu, mine.x, sl);
                        const int row = __float_as_int(__shfl_sync(0xffffffffu, mine.y, sl));
                        if (src[u] >= 0) load_row<VB, NCH>(raw[u], base + size_t(row) * kRowBytes);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (src[u] >= 0) {
#pragma unroll
                            for (int k = 0; k < NCH; ++k) fma_chunk<VB, kFp16>(acc[k], raw[u][k], cf[u]);
                        }
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
            // this block is the only writer of rows v0..v0+31: first pass stores, later passes accumulate
            float* out = dW + size_t(v) * (kRowBytes / 2) + lane * kPer;
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
#pragma unroll
                for (int i = 0; i < kPer; ++i) {
                    float* o = out + k * 32 * kPer + i;
                    *o = (b0 > 0 ? *o : 0.f) + acc[k][i];
                }
            }
            if (dbias != nullptr && lane == 0) dbias[v] = (b0 > 0 ? dbias[v] : 0.f) + bsum;
        }
    }
}

template <int VB, int NCH, bool kFp16>
__global__ void __launch_bounds__(256)
bwd_dh_fit_kernel(const uint2* __restrict__ entries, const int* __restrict__ nact, const char* __restrict__ W,
                  const int32_t* __restrict__ cu, int L, int V, float* __restrict__ d_hidden) {
    constexpr int kRowBytes = VB * 32 * NCH;
    constexpr int kPer = VB / 2;
    const int b = blockIdx.y;
    const int n = __ldg(nact + b);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int first = blockIdx.x * kDhEntriesPerBlock + warp * 32;
    if (first >= n) return;
    const int cnt = min(32, n - first);
    uint2 mine = make_uint2(0xffffffffu, 0u);       // lanes past the end: a key that starts no run of its own
    if (lane < cnt) mine = __ldg(entries + size_t(b) * V + first + lane);
    const uint32_t my_l = mine.x >> 20;
    // run structure of the (l-sorted) entries: bit e set <=> entry e starts a new run
    const uint32_t prev_l = __shfl_up_sync(0xffffffffu, my_l, 1);
    uint32_t starts = __ballot_sync(0xffffffffu, lane < cnt && (lane == 0 || my_l != prev_l));
    const char* wl = W + lane * VB;
    const size_t row0 = (cu != nullptr) ? size_t(__ldg(cu + b)) : size_t(b) * L;   // packed [T, H] or padded [B, L, H]
    float* dh = d_hidden + row0 * (kRowBytes / 2) + lane * kPer;
    while (starts != 0u) {
        const int s = __ffs(starts) - 1;
        starts &= starts - 1;
        const int e_end = starts != 0u ? (__ffs(starts) - 1) : cnt;
        float acc[NCH][kPer];
#pragma unroll
        for (int k = 0; k < NCH; ++k)
#pragma unroll
            for (int i = 0; i < kPer; ++i) acc[k][i] = 0.f;
        int e = s;
        for (; e + 4 <= e_end; e += 4) {
            typename Chunk<VB>::type raw[4][NCH];
            float cf[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t key = __shfl_sync(0xffffffffu, mine.x, e + u);
                cf[u] = __uint_as_float(__shfl_sync(0xffffffffu, mine.y, e + u));
                load_row<VB, NCH>(raw[u], wl + size_t(key & 0xFFFFFu) * kRowBytes);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < NCH; ++k) fma_chunk<VB, kFp16>(acc[k], raw[u][k], cf[u]);
        }
        for (; e < e_end; ++e) {
            const uint32_t key = __shfl_sync(0xffffffffu, mine.x, e);
            const float cf = __uint_as_float(__shfl_sync(0xffffffffu, mine.y, e));
            typename Chunk<VB>::type raw[NCH];
            load_row<VB, NCH>(raw, wl + size_t(key & 0xFFFFFu) * kRowBytes);
#pragma unroll
            for (int k = 0; k < NCH; ++k) fma_chunk<VB, kFp16>(acc[k], raw[k], cf);
        }
        const uint32_t l = __shfl_sync(0xffffffffu, my_l, s);
        float* row = dh + size_t(l) * (kRowBytes / 2);
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            float* o = row + k * 32 * kPer;
            if (kPer == 8) {
                red_add_v4(o, acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
                red_add_v4(o + 4, acc[k][4 % kPer], acc[k][5 % kPer], acc[k][6 % kPer], acc[k][7 % kPer]);
            } else if (kPer == 4) {
                red_add_v4(o, acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
            } else {
                asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(o), "f"(acc[k][0]), "f"(acc[k][1]) : "memory");
            }
        }
    }
}

// rep[b,v] *= (rep[b,v] > ratio * rowmax)   (sparse_encoders.py:118-119)
__global__ void __launch_bounds__(256) prune_rows_kernel(float* __restrict__ rep, int V, float ratio) {
    __shared__ float red[8];
    float* row = rep + size_t(blockIdx.x) * V;
    float m = -INFINITY;
    for (int v = threadIdx.x; v < V; v += 256) m = fmaxf(m, row[v]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    const float thr = m * ratio;
    for (int v = threadIdx.x; v < V; v += 256) {
        const float x = row[v];
        row[v] = (x > thr) ? x : x * 0.f;
    }
}

// exact-fit dispatch: returns false when H has no (VB, NCH) tiling
template <bool kFp16>
bool launch_bwd_fit(const float* d_rep, const float* xmax, const int32_t* argmax, const void* hidden, const void* W,
                    const int32_t* cu, size_t total_rows, int B, int L, int H, int V, int l0, float* d_hidden, float* dW,
                    float* dbias, uint2* entries, int* nact, cudaStream_t stream, int* rc) {
    int Bc = B;
    const int max_smem = 96 * 1024;
    if (size_t(Bc + 1) * kDwRows * sizeof(float2) > size_t(max_smem)) Bc = max_smem / int(kDwRows * sizeof(float2)) - 1;
    const size_t smem = size_t(Bc + 1) * kDwRows * sizeof(float2);
    const dim3 dh_grid((V + kDhEntriesPerBlock - 1) / kDhEntriesPerBlock, B);
    const int dw_grid = (V + kDwRows - 1) / kDwRows;
    const char* h = static_cast<const char*>(hidden);
    const char* w = static_cast<const char*>(W);
    *rc = SB200_OK;
#define SB200_FIT(VB, NCH, SLOT)                                                                                        \
    do {                                                                                                                \
        if (smem > 48 * 1024 && !device_flag_test_and_set(SLOT + (kFp16 ? 1 : 0))) {                                    \
            if (cudaFuncSetAttribute(bwd_dw_fit_kernel<VB, NCH, kFp16>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                     max_smem) != cudaSuccess) { *rc = fail(SB200_ERR_CUDA, "bwd_dw_fit: smem attribute"); return true; } \
        }                                                                                                               \
        bwd_dw_fit_kernel<VB, NCH, kFp16><<<dw_grid, kDwThreads, smem, stream>>>(d_rep, xmax, argmax, h, cu, B, L, V, l0, \
                                                                                 Bc, dW, dbias);                       \
        if (cudaGetLastError() != cudaSuccess) { *rc = fail(SB200_ERR_CUDA, "bwd_dw_fit_kernel launch"); return true; } \
        count_launch();                                                                                                 \
        if (cudaMemsetAsync(d_hidden, 0, total_rows * H * sizeof(float), stream) != cudaSuccess) {                      \
            *rc = fail(SB200_ERR_CUDA, "bwd: memset"); return true; }                                                   \
        bwd_bucket_kernel<<<B, kBucketThreads, 0, stream>>>(d_rep, xmax, argmax, L, V, l0, entries, nact);             \
        count_launch();                                                                                                 \
        bwd_dh_fit_kernel<VB, NCH, kFp16><<<dh_grid, 256, 0, stream>>>(entries, nact, w, cu, L, V, d_hidden);          \
        if (cudaGetLastError() != cudaSuccess) { *rc = fail(SB200_ERR_CUDA, "bwd_dh_fit_kernel launch"); return true; } \
        count_launch();                                                                                                 \
        return true;                                                                                                    \
    } while (0)
    switch (H) {
        case 64: SB200_FIT(4, 1, 20);
        case 128: SB200_FIT(8, 1, 22);
        case 256: SB200_FIT(16, 1, 24);
        case 384: SB200_FIT(8, 3, 26);
        case 512: SB200_FIT(16, 2, 28);
        case 768: SB200_FIT(16, 3, 30);
        default: break;
    }
#undef SB200_FIT
    return false;
}

template <int NCHUNK, bool kFp16>
int launch_bwd(const float* d_rep, const float* xmax, const int32_t* argmax, const __nv_bfloat16* hidden,
               const __nv_bfloat16* W, int B, int L, int H, int V, int l0, float* d_hidden, float* dW, float* dbias,
               uint2* entries, int* nact, cudaStream_t stream) {
    // dW / dbias
    {
        int Bc = B;
        const int max_smem = 96 * 1024;
        if (size_t(Bc) * kDwRows * sizeof(float2) > size_t(max_smem)) Bc = max_smem / int(kDwRows * sizeof(float2));
        const size_t smem = size_t(Bc) * kDwRows * sizeof(float2);
        if (!device_flag_test_and_set(kFp16 ? 15 + NCHUNK : NCHUNK))
            SB200_CUDA(cudaFuncSetAttribute(bwd_dw_kernel<NCHUNK, kFp16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            max_smem));
        bwd_dw_kernel<NCHUNK, kFp16><<<(V + kDwRows - 1) / kDwRows, kDwThreads, smem, stream>>>(
            d_rep, xmax, argmax, hidden, B, L, H, V, l0, Bc, dW, dbias);
        SB200_CHECK_LAUNCH("bwd_dw_kernel");
    }
    // d_hidden
    SB200_CUDA(cudaMemsetAsync(d_hidden, 0, size_t(B) * L * H * sizeof(float), stream));
    bwd_bucket_kernel<<<B, kBucketThreads, 0, stream>>>(d_rep, xmax, argmax, L, V, l0, entries, nact);
    SB200_CHECK_LAUNCH("bwd_bucket_kernel");
    dim3 grid((V + kDhEntriesPerBlock - 1) / kDhEntriesPerBlock, B);
    bwd_dh_kernel<NCHUNK, kFp16><<<grid, 256, 0, stream>>>(entries, nact, W, L, H, V, d_hidden);
    SB200_CHECK_LAUNCH("bwd_dh_kernel");
    return SB200_OK;
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" size_t sb200_head_bwd_workspace_bytes(int B, int L, int H, int V) {
    (void)L;
    (void)H;
    if (B <= 0 || V <= 0) return 0;
    return align_up(size_t(B) * V * sizeof(uint2), 256) + align_up(size_t(B) * sizeof(int), 256);
}

static int head_bwd_impl(const float* d_rep, const float* xmax, const int32_t* argmax, const void* hidden, const void* W,
                         const int32_t* cu, int T, int B, int L, int H, int V, int flags, float* d_hidden, float* dW,
                         float* dbias, void* workspace, size_t workspace_bytes, sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t total_rows = cu != nullptr ? size_t(T) : size_t(B) * L;
    SB200_REQUIRE(d_rep && xmax && argmax && hidden && W && d_hidden && dW, "head_bwd: null pointer");
    SB200_REQUIRE(B >= 1 && L >= 1 && L <= kMaxL && V >= 1 && V <= (1 << 20), "head_bwd: bad shape B=%d L=%d V=%d", B,
                  L, V);
    SB200_REQUIRE(H >= 8 && H % 8 == 0 && H <= 256 * kMaxChunks, "head_bwd: H=%d must be a multiple of 8, <= %d", H,
                  256 * kMaxChunks);
    SB200_REQUIRE(B <= 65535, "head_bwd: B=%d exceeds grid.y", B);
    const size_t need = sb200_head_bwd_workspace_bytes(B, L, H, V);
    if (workspace == nullptr || workspace_bytes < need)
        return fail(SB200_ERR_WORKSPACE, "head_bwd: workspace %zu < %zu", workspace_bytes, need);
    uint2* entries = static_cast<uint2*>(workspace);
    int* nact = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + align_up(size_t(B) * V * sizeof(uint2), 256));
    const int l0 = (flags & SB200_HEAD_L0) ? 1 : 0;
    const __nv_bfloat16* h = static_cast<const __nv_bfloat16*>(hidden);
    const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(W);
    {
        int rc = SB200_OK;
        const bool done = (flags & SB200_HEAD_FP16)
            ? launch_bwd_fit<true>(d_rep, xmax, argmax, hidden, W, cu, total_rows, B, L, H, V, l0, d_hidden, dW, dbias, entries, nact, stream, &rc)
            : launch_bwd_fit<false>(d_rep, xmax, argmax, hidden, W, cu, total_rows, B, L, H, V, l0, d_hidden, dW, dbias, entries, nact, stream, &rc);
        if (done) return rc;
        if (cu != nullptr) return fail(SB200_ERR_ARG, "head_bwd_packed: H=%d has no exact-fit kernel (64/128/256/384/512/768)", H);
    }
    const int nchunk = (H / 8 + 31) / 32;
#define SB200_BWD(N, F) launch_bwd<N, F>(d_rep, xmax, argmax, h, w, B, L, H, V, l0, d_hidden, dW, dbias, entries, nact, stream)
    if (flags & SB200_HEAD_FP16) {   // the 16-bit payload is reinterpreted inside the kernels
        switch (nchunk) {
            case 1: return SB200_BWD(1, true);
            case 2: return SB200_BWD(2, true);
            case 3: return SB200_BWD(3, true);
            default: return SB200_BWD(4, true);
        }
    }
    switch (nchunk) {
        case 1: return SB200_BWD(1, false);
        case 2: return SB200_BWD(2, false);
        case 3: return SB200_BWD(3, false);
        default: return SB200_BWD(4, false);
    }
#undef SB200_BWD
}

extern "C" int sb200_head_bwd(const float* d_rep, const float* xmax, const int32_t* argmax, const void* hidden,
                              const void* W, int B, int L, int H, int V, int flags, float* d_hidden, float* dW,
                              float* dbias, void* workspace, size_t workspace_bytes, sb200_stream_t stream) {
    return head_bwd_impl(d_rep, xmax, argmax, hidden, W, nullptr, 0, B, L, H, V, flags, d_hidden, dW, dbias, workspace,
                         workspace_bytes, stream);
}

extern "C" int sb200_head_bwd_packed(const float* d_rep, const float* xmax, const int32_t* argmax, const void* hidden,
                                     const void* W, const int32_t* cu_seqlens, int T, int B, int max_len, int H, int V,
                                     int flags, float* d_hidden, float* dW, float* dbias, void* workspace,
                                     size_t workspace_bytes, sb200_stream_t stream) {
    if (cu_seqlens == nullptr || T < 1) return fail(SB200_ERR_ARG, "head_bwd_packed: cu_seqlens / T");
    return head_bwd_impl(d_rep, xmax, argmax, hidden, W, cu_seqlens, T, B, max_len, H, V, flags, d_hidden, dW, dbias,
                         workspace, workspace_bytes, stream);
}

extern "C" int sb200_head_packed_supported(int H, int max_len) {
    return (max_len > 128 && max_len <= 4096 && (H == 64 || H == 128 || H == 256 || H == 384 || H == 512 || H == 768)) ? 1 : 0;
}

extern "C" int sb200_prune_rows(float* rep, int B, int V, float ratio, sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SB200_REQUIRE(rep && B >= 1 && V >= 1, "prune_rows: bad arguments");
    prune_rows_kernel<<<B, 256, 0, stream>>>(rep, V, ratio);
    SB200_CHECK_LAUNCH("prune_rows_kernel");
    return SB200_OK;
}
