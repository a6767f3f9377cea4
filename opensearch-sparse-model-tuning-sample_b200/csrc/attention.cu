// Variable-length (padding-free) multi-head self-attention of the BERT body, forward and backward, for head_dim 32 / 64:
//   P = softmax(Q K^T * scale) per (sequence, head), O = dropout(P) V
// (transformers BertSelfAttention.forward inside the backbone call at scripts/model/sparse_encoders.py:108, and its
// autograd chain). Sequences are packed back to back in one [T, .] token matrix and delimited by cu_seqlens, so no
// padding row is ever multiplied.
//
// Why warp-level mma.sync and not tcgen05 here: at head_dim 32 a score element costs 4 tensor flops per exp2 / max /
// sum / convert / dropout decision -- the kernel is bound by the FP32 + MUFU pipes and (through TMEM) by the same
// 64 B/clk read port that bounds the head epilogue, not by the tensor pipe. With mma.sync the score fragment IS the
// A fragment of the following P V product, so S and P never leave the register file. See DESIGN.md 4.6.
//
// Work split: one CTA = one (sequence, head, tile of 64 rows); 4 warps x 16 rows; the other operand streams through a
// 2-stage cp.async ring of 64-row blocks (XOR-swizzled, conflict-free for ldmatrix).
//   forward   warp owns 16 queries: S = Q K^T (64 keys) -> online softmax -> dropout -> O += P V.   Saves LSE.
//   backward  two CTA roles in ONE launch, no atomics, no cross-warp reduction, every output written exactly once:
//     role dQ     warp owns 16 queries, streams K/V:  S, P, dP = dO V^T, dS = P (dP - D)  ->  dQ += dS K
//     role dK/dV  warp owns 16 keys, streams Q/dO (transposed problem S^T = K Q^T): dV += P^T dO, dK += dS^T Q
//   D = rowsum(dO * O) comes from a small pre-pass.
// Dropout: the keep mask is a pure function of (seed, salt, sequence, head, query, key): one 32-bit hash per 2x2 patch
// {r, r+8} x {c, c+8} of a 16x16 score block, a byte per element. That patch is exactly what one thread holds in the
// m16n8k16 accumulator layout in BOTH orientations (queries x keys and keys x queries), so the forward and both
// backward roles regenerate identical masks from registers. Keep probability is quantised to thr/256 and the
// rescale uses the quantised value (unbiased).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "common.h"
#include "ptx.cuh"

namespace sb200 {
namespace {

constexpr int kAttThreads = 128;
// resident CTAs per SM the register allocation is capped for (head_dim 32 / 64)
#ifndef SB200_ATT_FWD_CTAS32
#define SB200_ATT_FWD_CTAS32 4
#define SB200_ATT_FWD_CTAS64 3
#define SB200_ATT_BWD_CTAS32 4
#define SB200_ATT_BWD_CTAS64 3
#endif
// 1: separate code for full 64-row blocks and tail blocks in the backward roles; 0: one (tail-capable) path, half the code
#ifndef SB200_ATT_BWD_SPLIT
#define SB200_ATT_BWD_SPLIT 1
#endif
// 1: the two backward roles as two launches (half the code per kernel); 0: one launch
#ifndef SB200_ATT_BWD_TWO_LAUNCHES
#define SB200_ATT_BWD_TWO_LAUNCHES 0
#endif
constexpr int kAttTile = 64;
// Stages of the operand ring: 4 x 64 rows cover sequences up to 256 tokens without ever reusing a stage (no block-wide
// barrier in the loop at all); longer sequences refill a stage behind one __syncthreads.
template <int D>
constexpr int att_stages() { return D == 32 ? 4 : 3; }
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct AttnParams {
    const __nv_bfloat16* q;   // [T, .] row stride in_stride, this head at + head * D
    const __nv_bfloat16* k;
    const __nv_bfloat16* v;
    long long in_stride;
    __nv_bfloat16* out;       // [T, h * D]
    float* lse;               // [h, T] natural-log LSE of the scaled scores
    const int* cu;            // [nseq + 1]
    int T, h, nseq, ntile;
    int nseq_live;            // sequences [nseq_live, nseq) are filler: their output rows are zero-filled, nothing is computed
    float scale, scale_log2;
    uint32_t keep_thr;        // keep probability = keep_thr / 256  (256 = keep everything)
    uint32_t keep_add;        // (128 - (256 - keep_thr)) in every byte: SWAR compare byte >= 256 - keep_thr
    float inv_keep;
    const unsigned long long* seed;
    uint32_t salt;
    // backward
    const __nv_bfloat16* dout;  // [T, h * D]
    const float* stat;          // [2, h, T]: lse * log2e - log2(inv_keep), rowsum(dout * out) / inv_keep
    __nv_bfloat16* dq;          // same layout as q / k / v (row stride d_stride)
    __nv_bfloat16* dk;
    __nv_bfloat16* dv;
    long long d_stride;
    unsigned char* mask_out;    // test hook (dropout mask dump)
};

// ---------------------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// Makes `bar` track the completion of every cp.async this thread has issued so far (counted in the barrier's
// initial arrival count: one such arrival per thread and phase).
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

// ---------------------------------------------------------------------------------------------------------- tiles
// A [64][D] bf16 tile in shared memory; 16-byte chunks XOR-swizzled so that the 8 row addresses of every ldmatrix
// 8x8 matrix fall into 8 different bank groups (D = 32: rows are 64 B, two rows per 128 B line).
template <int D>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
    if (D == 32) return uint32_t(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
    return uint32_t(row * 128 + ((chunk ^ (row & 7)) << 4));
}

// Per-thread state of a stream of 64-row blocks of one [rows, D] operand of one sequence (rows >= len zero-filled):
// everything that does not change from block to block (row / chunk of the thread, swizzled offset, strides) is
// computed once, so that staging a block costs a handful of instructions per thread.
template <int D>
struct TileStream {
    static constexpr int CPR = D / 8;                   // 16-byte chunks per row
    static constexpr int RPP = kAttThreads / CPR;       // rows per pass (the swizzle repeats every 8 rows)
    const __nv_bfloat16* base;                          // a valid address for the zero-fill form
    const __nv_bfloat16* next;
    long long step, sub;
    int row;
    uint32_t dst_off;
    __device__ __forceinline__ void init(const __nv_bfloat16* seq_base, long long stride, int row0, int tid) {
        const int r = tid / CPR, c = tid % CPR;
        base = seq_base;
        step = kAttTile * stride;
        sub = RPP * stride;
        row = row0 + r;
        next = seq_base + (long long)row * stride + c * 8;
        dst_off = tile_off<D>(r, c);
    }
    __device__ __forceinline__ void load(uint32_t tile, int len) {   // stages the next block and advances
#pragma unroll
        for (int i = 0; i < kAttTile / RPP; ++i) {
            const bool ok = row + i * RPP < len;
            cp_async16(tile + dst_off + i * RPP * D * 2, ok ? next + i * sub : base, ok ? 16 : 0);
        }
        next += step;
        row += kAttTile;
    }
};

// Same for a per-row fp32 statistic (threads 0..63, one float each; rows >= len -> 0).
struct StatStream {
    const float* base;
    const float* next;
    int row;
    uint32_t dst_off;
    __device__ __forceinline__ void init(const float* seq_base, int tid) {
        base = seq_base;
        row = tid;
        next = seq_base + tid;
        dst_off = uint32_t(tid) * 4u;
    }
    __device__ __forceinline__ void load(uint32_t dst, int len) {
        if (dst_off < kAttTile * 4u) {
            const bool ok = row < len;
            cp_async4(dst + dst_off, ok ? next : base, ok ? 4 : 0);
        }
        next += kAttTile;
        row += kAttTile;
    }
};

// "A pattern": the 16 x 16 region at (row0, 16 * kc): r0 = rows 0-7 / cols 0-7, r1 = rows 8-15 / cols 0-7,
// r2 = rows 0-7 / cols 8-15, r3 = rows 8-15 / cols 8-15. Plain: the A fragment. Transposed (stored rows = k index):
// {r0, r1} = B fragment of n-tile 2 * kc, {r2, r3} = B fragment of n-tile 2 * kc + 1.
template <int D>
__device__ __forceinline__ uint32_t addr_a(uint32_t tile, int row0, int kc, int lane) {
    return tile + tile_off<D>(row0 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * kc + (lane >> 4));
}
// "B pattern": 8 stored rows (n index) x 32 columns (k index) starting at chunk c0: r[2j], r[2j+1] = B fragment of
// k-step (c0 / 2 + j).
template <int D>
__device__ __forceinline__ uint32_t addr_b(uint32_t tile, int row0, int c0, int lane) {
    return tile + tile_off<D>(row0 + (lane & 7), c0 + (lane >> 3));
}

// Loads the A fragments (all k-steps) of the warp's 16 rows.
template <int D>
__device__ __forceinline__ void load_a_frags(uint32_t tile, int row0, int lane, uint32_t (&a)[D / 16][4]) {
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) ldsm_x4(addr_a<D>(tile, row0, ks, lane), a[ks]);
}

// acc[nt] (16 x 8) += A[16 x D] * B^T where B = 8 stored rows starting at brow0 of `tile`
template <int D>
__device__ __forceinline__ void mma_rows(float (&acc)[4], const uint32_t (&a)[D / 16][4], uint32_t tile, int brow0,
                                         int lane) {
#pragma unroll
    for (int gi = 0; gi < D / 32; ++gi) {
        uint32_t b[4];
        ldsm_x4(addr_b<D>(tile, brow0, 4 * gi, lane), b);
        mma_bf16(acc, a[2 * gi], b[0], b[1]);
        mma_bf16(acc, a[2 * gi + 1], b[2], b[3]);
    }
}

// acc[D / 8] (16 x D) += A (16 x 16, registers) * tile rows [krow0, krow0 + 16) (stored rows = k index)
template <int D>
__device__ __forceinline__ void mma_cols(float (&acc)[D / 8][4], const uint32_t (&a)[4], uint32_t tile, int krow0,
                                         int lane) {
#pragma unroll
    for (int np = 0; np < D / 16; ++np) {
        uint32_t b[4];
        ldsm_x4_t(addr_a<D>(tile, krow0, np, lane), b);
        mma_bf16(acc[2 * np], a, b[0], b[1]);
        mma_bf16(acc[2 * np + 1], a, b[2], b[3]);
    }
}

// Writes the warp's 16 x D accumulator (scaled) as bf16 through its own 16 rows of `tile` to global rows.
template <int D>
__device__ __forceinline__ void store_rows(const float (&acc)[D / 8][4], float s0, float s1,
                                           unsigned char* tile_ptr, int wrow0, __nv_bfloat16* dst, long long stride,
                                           int grow0, int len, int lane) {
    constexpr int CPR = D / 8;
    const int g = lane >> 2, t = lane & 3;
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
        *reinterpret_cast<uint32_t*>(tile_ptr + tile_off<D>(wrow0 + g, nt) + t * 4) =
            pack_bf16(acc[nt][0] * s0, acc[nt][1] * s0);
        *reinterpret_cast<uint32_t*>(tile_ptr + tile_off<D>(wrow0 + g + 8, nt) + t * 4) =
            pack_bf16(acc[nt][2] * s1, acc[nt][3] * s1);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16 * CPR / 32; ++i) {
        const int idx = lane + i * 32;
        const int r = idx / CPR, c = idx % CPR;
        const int gr = grow0 + wrow0 + r;
        if (gr < len) {
            const uint4 val = *reinterpret_cast<const uint4*>(tile_ptr + tile_off<D>(wrow0 + r, c));
            *reinterpret_cast<uint4*>(dst + (long long)gr * stride + c * 8) = val;
        }
    }
}

// Zero-fills the D columns of one head in rows [row0, row0 + 64) of a sequence (filler sequences of a packed batch).
template <int D>
__device__ __forceinline__ void zero_rows(__nv_bfloat16* dst, long long stride, int row0, int len, int tid) {
    constexpr int CPR = D / 8;
#pragma unroll
    for (int i = 0; i < kAttTile * CPR / kAttThreads; ++i) {
        const int idx = tid + i * kAttThreads;
        const int r = row0 + idx / CPR, c = idx % CPR;
        if (r < len) *reinterpret_cast<uint4*>(dst + (long long)r * stride + c * 8) = make_uint4(0u, 0u, 0u, 0u);
    }
}

__device__ __forceinline__ uint32_t seq_key(const AttnParams& p, int seq, int head) {
    const unsigned long long s = *p.seed;
    return mix32(uint32_t(s) ^ mix32(uint32_t(s >> 32) + 0x9E3779B9u * uint32_t(seq * p.h + head + 1)) ^
                 (p.salt * 0x85EBCA6Bu));
}
// Dropout decisions of one 2x2 patch {r, r+8} x {c, c+8} of the 16 x 16 score block (qblk, kblk): a 32-bit hash, one
// byte per element (byte 0: (r, c), 1: (r, c+8), 2: (r+8, c), 3: (r+8, c+8)); an element is kept iff its byte is
// >= 256 - thr. Returned as flags: bit 7 of byte i set <=> element i kept (SWAR compare; needs 256 - thr <= 128).
constexpr uint32_t kGolden = 0x9E3779B1u;
__device__ __forceinline__ uint32_t patch_index(int qblk, int kblk, int r, int c) {
    return uint32_t(((qblk * 64 + kblk) * 64) + r * 8 + c);
}
__device__ __forceinline__ uint32_t patch_flags(uint32_t key, uint32_t idx_times_golden, uint32_t addc) {
    // multiply-fold ("mum") of the keyed, golden-ratio-scrambled patch index: high ^ low word of a 32 x 32 -> 64 product
    const unsigned long long m = (unsigned long long)(key ^ idx_times_golden) * 0xD6E8FEB9u;
    const uint32_t h = uint32_t(m >> 32) ^ uint32_t(m);
    return ((h & 0x7F7F7F7Fu) + addc) | h;
}
template <uint32_t kSel>
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "n"(kSel));
    return d;
}
// bf16x2 mask of byte I of two patches: low half from f0, high half from f1 (selector bit 3 = replicate the byte's msb)
template <int I>
__device__ __forceinline__ uint32_t pair_mask(uint32_t f0, uint32_t f1) {
    return prmt<0xCC88u + 0x1111u * I>(f0, f1);
}
template <int I>
__device__ __forceinline__ float mask_f32(float x, uint32_t f) {
    return __uint_as_float(__float_as_uint(x) & prmt<0x8888u + 0x1111u * I>(f, f));
}

// ---------------------------------------------------------------------------------------------------------- forward
// One block of the forward pass for the warp's 16 queries: NC chunks of 16 keys (NC = 4: a full 64-key block).
// kMask: the last chunk may reach beyond the sequence (keys >= lim are masked); blocks are specialised by their
// number of valid chunks so that a short tail block costs what its keys cost.
template <int D, bool kDrop, int NC, bool kMask>
__device__ __forceinline__ void fwd_block(const AttnParams& p, const uint32_t (&qa)[D / 16][4], float (&o)[D / 8][4],
                                          float& m0, float& m1, float& l0, float& l1, uint32_t tK, uint32_t tV, int lim,
                                          uint32_t key, uint32_t ig, int lane) {
    const int t = lane & 3;
    float s[2 * NC][4];
#pragma unroll
    for (int nt = 0; nt < 2 * NC; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        mma_rows<D>(s[nt], qa, tK, nt * 8, lane);
    }
    if (kMask) {
#pragma unroll
        for (int nt = 2 * NC - 2; nt < 2 * NC; ++nt) {
            const int c = nt * 8 + 2 * t;
            if (c >= lim) s[nt][0] = s[nt][2] = -INFINITY;
            if (c + 1 >= lim) s[nt][1] = s[nt][3] = -INFINITY;
        }
    }
    float mx0 = s[0][0], mx1 = s[0][2];
#pragma unroll
    for (int nt = 0; nt < 2 * NC; ++nt) {
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = ex2((m0 - mn0) * p.scale_log2), c1 = ex2((m1 - mn1) * p.scale_log2);
    m0 = mn0;
    m1 = mn1;
    const float b0 = mn0 * p.scale_log2, b1 = mn1 * p.scale_log2;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2 * NC; ++nt) {
        s[nt][0] = ex2(fmaf(s[nt][0], p.scale_log2, -b0));
        s[nt][1] = ex2(fmaf(s[nt][1], p.scale_log2, -b0));
        s[nt][2] = ex2(fmaf(s[nt][2], p.scale_log2, -b1));
        s[nt][3] = ex2(fmaf(s[nt][3], p.scale_log2, -b1));
        sum0 += s[nt][0] + s[nt][1];
        sum1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * c0 + sum0;
    l1 = l1 * c1 + sum1;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        o[i][0] *= c0;
        o[i][1] *= c0;
        o[i][2] *= c1;
        o[i][3] *= c1;
    }
#pragma unroll
    for (int kk = 0; kk < NC; ++kk) {
        uint32_t a[4];
        a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
        a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
        a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        if (kDrop) {
            const uint32_t igk = ig + uint32_t(kk * 64) * kGolden;
            const uint32_t f0 = patch_flags(key, igk, p.keep_add), f1 = patch_flags(key, igk + kGolden, p.keep_add);
            a[0] &= pair_mask<0>(f0, f1);
            a[1] &= pair_mask<2>(f0, f1);
            a[2] &= pair_mask<1>(f0, f1);
            a[3] &= pair_mask<3>(f0, f1);
        }
        mma_cols<D>(o, a, tV, kk * 16, lane);
    }
}

template <int D, bool kDrop>
__global__ void __launch_bounds__(kAttThreads, D == 32 ? SB200_ATT_FWD_CTAS32 : SB200_ATT_FWD_CTAS64) attn_fwd_kernel(const AttnParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int kTileBytes = kAttTile * D * 2;
    const int seq = blockIdx.z, head = blockIdx.y, qt = blockIdx.x;
    const int s0 = __ldg(p.cu + seq);
    const int len = __ldg(p.cu + seq + 1) - s0;
    if (qt * kAttTile >= len) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    if (seq >= p.nseq_live) {
        zero_rows<D>(p.out + (long long)s0 * (p.h * D) + head * D, (long long)p.h * D, qt * kAttTile, len, tid);
        if (tid < kAttTile && qt * kAttTile + tid < len) p.lse[(long long)head * p.T + s0 + qt * kAttTile + tid] = 0.f;
        return;
    }
    constexpr int NS = att_stages<D>();
    const uint32_t sQ = smem_u32(smem), sK = sQ + kTileBytes, sV = sK + NS * kTileBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (1 + 2 * NS) * kTileBytes);
    const __nv_bfloat16* qs = p.q + (long long)s0 * p.in_stride + head * D;
    const __nv_bfloat16* ks = p.k + (long long)s0 * p.in_stride + head * D;
    const __nv_bfloat16* vs = p.v + (long long)s0 * p.in_stride + head * D;
    const int nb = (len + kAttTile - 1) / kAttTile;
    const bool active = qt * kAttTile + warp * 16 < len;   // warps whose 16 rows are all padding only load

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) mbar_init(full + i, kAttThreads);
        fence_mbar_init();
    }
    __syncthreads();
    TileStream<D> tq, tk, tv;
    tq.init(qs, p.in_stride, qt * kAttTile, tid);
    tk.init(ks, p.in_stride, 0, tid);
    tv.init(vs, p.in_stride, 0, tid);
    tq.load(sQ, len);
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        if (i < nb) {
            tk.load(sK + i * kTileBytes, len);
            tv.load(sV + i * kTileBytes, len);
            cp_async_arrive(full + i);
        }
    }

    uint32_t qa[D / 16][4];
    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    uint32_t key = 0;
    if (kDrop) key = seq_key(p, seq, head);
    const uint32_t ig0 = patch_index(qt * 4 + warp, 0, g, 2 * t) * kGolden;

    for (int kb = 0; kb < nb; ++kb) {
        const int st = kb % NS;
        if (active) {
            mbar_wait(full + st, (kb / NS) & 1);
            if (kb == 0) load_a_frags<D>(sQ, warp * 16, lane, qa);
            const uint32_t tK = sK + st * kTileBytes, tV = sV + st * kTileBytes;
            const int lim = len - kb * kAttTile;     // valid keys in this block
            const uint32_t ig = ig0 + uint32_t(kb * 4 * 64) * kGolden;
            if (lim >= kAttTile)
                fwd_block<D, kDrop, 4, false>(p, qa, o, m0, m1, l0, l1, tK, tV, lim, key, ig, lane);
            else if (lim > 48)
                fwd_block<D, kDrop, 4, true>(p, qa, o, m0, m1, l0, l1, tK, tV, lim, key, ig, lane);
            else if (lim > 32)
                fwd_block<D, kDrop, 3, true>(p, qa, o, m0, m1, l0, l1, tK, tV, lim, key, ig, lane);
            else if (lim > 16)
                fwd_block<D, kDrop, 2, true>(p, qa, o, m0, m1, l0, l1, tK, tV, lim, key, ig, lane);
            else
                fwd_block<D, kDrop, 1, true>(p, qa, o, m0, m1, l0, l1, tK, tV, lim, key, ig, lane);
        }
        if (kb + NS < nb) {     // long sequence: the stage is reused once every warp is done with it
            __syncthreads();
            tk.load(sK + st * kTileBytes, len);
            tv.load(sV + st * kTileBytes, len);
            cp_async_arrive(full + st);
        }
    }
    if (!active) return;
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = p.inv_keep / l0, i1 = p.inv_keep / l1;
    store_rows<D>(o, i0, i1, smem, warp * 16, p.out + (long long)s0 * (p.h * D) + head * D, (long long)p.h * D,
                  qt * kAttTile, len, lane);
    if (t == 0) {
        const int r0 = qt * kAttTile + warp * 16 + g;
        float* lse = p.lse + (long long)head * p.T + s0;
        if (r0 < len) lse[r0] = (m0 * p.scale_log2 + log2f(l0)) * kLn2;
        if (r0 + 8 < len) lse[r0 + 8] = (m1 * p.scale_log2 + log2f(l1)) * kLn2;
    }
}

// ---------------------------------------------------------------------------------------------------------- pre-pass
// Per (head, token) constants of the backward pass; one warp per token row:
//   stat[0][head, t] = lse * log2(e) - log2(inv_keep)        (so that exp2(s * scale_log2 - .) = P * inv_keep)
//   stat[1][head, t] = rowsum(dout * out) / inv_keep          (D / inv_keep)
template <int D>
__global__ void __launch_bounds__(256) attn_prep_kernel(const __nv_bfloat16* __restrict__ dout,
                                                        const __nv_bfloat16* __restrict__ out,
                                                        const float* __restrict__ lse, int T, int h, float inv_keep,
                                                        float* __restrict__ stat) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= T) return;
    const int H = h * D;
    const float lg = log2f(inv_keep), rk = 1.f / inv_keep;
    for (int c0 = 0; c0 < H; c0 += 256) {   // uniform trip count: the shuffles below need the whole warp
        const int c = c0 + lane * 8;
        float acc = 0.f;
        if (c < H) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(dout + (long long)row * H + c));
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(out + (long long)row * H + c));
            const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
            const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 fa = __bfloat1622float2(pa[i]), fb = __bfloat1622float2(pb[i]);
                acc = fmaf(fa.x, fb.x, acc);
                acc = fmaf(fa.y, fb.y, acc);
            }
        }
#pragma unroll
        for (int off = 1; off < D / 8; off <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (c < H && (lane & (D / 8 - 1)) == 0) {
            const long long at = (long long)(c / D) * T + row;
            stat[at] = fmaf(__ldg(lse + at), kLog2e, -lg);
            stat[(long long)h * T + at] = acc * rk;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------- backward
// ds = P' * (mask * dp - D') for one accumulator fragment pair (16 x 16 chunk), with P' = exp2(s * scale_log2 - l2').
// Role dQ: rows = queries (l2 / dd per row: x = row g, y = row g + 8). Role dK/dV: rows = keys, columns = queries
// (l2 / dd per column).
template <bool kDrop, bool kRowStats, int N2>
__device__ __forceinline__ void bwd_chunk_math(float (&s)[4], float (&dp)[4], float sl2, float2 l2, float2 dd, uint32_t f0,
                                               uint32_t f1) {
    // element j: accumulator row half (j >> 1), column parity (j & 1); dropout byte: role dQ -> N2 + (j & 2),
    // role dK/dV -> 2 * N2 + (j >> 1); hash f0 for even columns, f1 for odd ones
#define SB200_ATT_ELEM(J)                                                                              \
    {                                                                                                  \
        const float l = kRowStats ? ((J & 2) ? l2.y : l2.x) : ((J & 1) ? l2.y : l2.x);                 \
        const float d = kRowStats ? ((J & 2) ? dd.y : dd.x) : ((J & 1) ? dd.y : dd.x);                 \
        const float pr = ex2(fmaf(s[J], sl2, -l));                                                     \
        float dpe = dp[J];                                                                             \
        if (kDrop) dpe = mask_f32<kRowStats ? (N2 + (J & 2)) : (2 * N2 + (J >> 1))>(dpe, (J & 1) ? f1 : f0); \
        dp[J] = pr;                                                                                    \
        s[J] = pr * (dpe - d);                                                                         \
    }
    SB200_ATT_ELEM(0) SB200_ATT_ELEM(1) SB200_ATT_ELEM(2) SB200_ATT_ELEM(3)
#undef SB200_ATT_ELEM
}

// One 64-key block of role dQ (see attn_bwd_dq) in chunks of 16 keys.
template <int D, bool kDrop, bool kTail>
__device__ __forceinline__ void bwd_dq_block(const AttnParams& p, const uint32_t (&qa)[D / 16][4],
                                             const uint32_t (&da)[D / 16][4], float (&dq)[D / 8][4], float2 l2, float2 dd,
                                             uint32_t tK, uint32_t tV, int lim, uint32_t key, uint32_t ig, int lane) {
    const int t = lane & 3;
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
        if (!kTail || kc * 16 < lim) {
            float s[2][4], dp[2][4];
#pragma unroll
            for (int n2 = 0; n2 < 2; ++n2) {
                s[n2][0] = s[n2][1] = s[n2][2] = s[n2][3] = 0.f;
                dp[n2][0] = dp[n2][1] = dp[n2][2] = dp[n2][3] = 0.f;
                mma_rows<D>(s[n2], qa, tK, kc * 16 + n2 * 8, lane);
                mma_rows<D>(dp[n2], da, tV, kc * 16 + n2 * 8, lane);
            }
            if (kTail && kc * 16 + 16 > lim) {   // the chunk straddles the end of the sequence: P = 0 beyond it
#pragma unroll
                for (int n2 = 0; n2 < 2; ++n2) {
                    const int c = kc * 16 + n2 * 8 + 2 * t;
                    if (c >= lim) s[n2][0] = s[n2][2] = -INFINITY;
                    if (c + 1 >= lim) s[n2][1] = s[n2][3] = -INFINITY;
                }
            }
            uint32_t f0 = 0, f1 = 0;
            if (kDrop) {
                const uint32_t igk = ig + uint32_t(kc * 64) * kGolden;
                f0 = patch_flags(key, igk, p.keep_add);
                f1 = patch_flags(key, igk + kGolden, p.keep_add);
            }
            bwd_chunk_math<kDrop, true, 0>(s[0], dp[0], p.scale_log2, l2, dd, f0, f1);
            bwd_chunk_math<kDrop, true, 1>(s[1], dp[1], p.scale_log2, l2, dd, f0, f1);
            uint32_t a[4];
            a[0] = pack_bf16(s[0][0], s[0][1]);
            a[1] = pack_bf16(s[0][2], s[0][3]);
            a[2] = pack_bf16(s[1][0], s[1][1]);
            a[3] = pack_bf16(s[1][2], s[1][3]);
            mma_cols<D>(dq, a, tK, kc * 16, lane);
        }
    }
}

// Role dQ: the warp owns 16 queries (Q and dO fragments in registers) and streams K / V.
template <int D, bool kDrop>
__device__ __forceinline__ void attn_bwd_dq(const AttnParams& p, unsigned char* smem, int seq, int head, int qt, int s0,
                                            int len) {
    constexpr int kTileBytes = kAttTile * D * 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    constexpr int NS = att_stages<D>();
    const uint32_t sQ = smem_u32(smem), sDO = sQ + kTileBytes, sK = sDO + kTileBytes, sV = sK + NS * kTileBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (2 + 2 * NS) * kTileBytes + 2 * NS * kAttTile * 4);
    const __nv_bfloat16* qs = p.q + (long long)s0 * p.in_stride + head * D;
    const __nv_bfloat16* ks = p.k + (long long)s0 * p.in_stride + head * D;
    const __nv_bfloat16* vs = p.v + (long long)s0 * p.in_stride + head * D;
    const __nv_bfloat16* dos = p.dout + (long long)s0 * (p.h * D) + head * D;
    const int nb = (len + kAttTile - 1) / kAttTile;
    const bool active = qt * kAttTile + warp * 16 < len;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) mbar_init(full + i, kAttThreads);
        fence_mbar_init();
    }
    __syncthreads();
    TileStream<D> tq, tk, tv;
    tq.init(qs, p.in_stride, qt * kAttTile, tid);
    tq.load(sQ, len);
    tq.init(dos, (long long)p.h * D, qt * kAttTile, tid);
    tq.load(sDO, len);
    tk.init(ks, p.in_stride, 0, tid);
    tv.init(vs, p.in_stride, 0, tid);
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        if (i < nb) {
            tk.load(sK + i * kTileBytes, len);
            tv.load(sV + i * kTileBytes, len);
            cp_async_arrive(full + i);
        }
    }

    const int r0 = qt * kAttTile + warp * 16 + g;
    const float* st_l = p.stat + (long long)head * p.T + s0;
    const float* st_d = st_l + (long long)p.h * p.T;
    float2 l2, dd;
    l2.x = r0 < len ? __ldg(st_l + r0) : 0.f;
    l2.y = r0 + 8 < len ? __ldg(st_l + r0 + 8) : 0.f;
    dd.x = r0 < len ? __ldg(st_d + r0) : 0.f;
    dd.y = r0 + 8 < len ? __ldg(st_d + r0 + 8) : 0.f;

    uint32_t qa[D / 16][4], da[D / 16][4];
    float dq[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    uint32_t key = 0;
    if (kDrop) key = seq_key(p, seq, head);
    const uint32_t ig0 = patch_index(qt * 4 + warp, 0, g, 2 * t) * kGolden;

    for (int kb = 0; kb < nb; ++kb) {
        const int st = kb % NS;
        if (active) {
            mbar_wait(full + st, (kb / NS) & 1);
            if (kb == 0) {
                load_a_frags<D>(sQ, warp * 16, lane, qa);
                load_a_frags<D>(sDO, warp * 16, lane, da);
            }
            const uint32_t tK = sK + st * kTileBytes, tV = sV + st * kTileBytes;
            const int lim = len - kb * kAttTile;
            const uint32_t ig = ig0 + uint32_t(kb * 4 * 64) * kGolden;
            if (SB200_ATT_BWD_SPLIT && lim >= kAttTile)
                bwd_dq_block<D, kDrop, false>(p, qa, da, dq, l2, dd, tK, tV, lim, key, ig, lane);
            else
                bwd_dq_block<D, kDrop, true>(p, qa, da, dq, l2, dd, tK, tV, lim, key, ig, lane);
        }
        if (kb + NS < nb) {
            __syncthreads();
            tk.load(sK + st * kTileBytes, len);
            tv.load(sV + st * kTileBytes, len);
            cp_async_arrive(full + st);
        }
    }
    if (!active) return;
    store_rows<D>(dq, p.scale, p.scale, smem, warp * 16, p.dq + (long long)s0 * p.d_stride + head * D, p.d_stride,
                  qt * kAttTile, len, lane);
}

// One 64-query block of role dK/dV (see attn_bwd_dkv) in chunks of 16 queries.
template <int D, bool kDrop, bool kTail>
__device__ __forceinline__ void bwd_dkv_block(const AttnParams& p, const uint32_t (&ka)[D / 16][4],
                                              const uint32_t (&va)[D / 16][4], float (&dk)[D / 8][4],
                                              float (&dv)[D / 8][4], const float* Ls, const float* Ds, uint32_t tQ,
                                              uint32_t tDO, int lim, uint32_t key, uint32_t ig, int lane) {
    const int t = lane & 3;
#pragma unroll
    for (int qc = 0; qc < 4; ++qc) {
        if (!kTail || qc * 16 < lim) {
            float s[2][4], dp[2][4];
#pragma unroll
            for (int n2 = 0; n2 < 2; ++n2) {
                s[n2][0] = s[n2][1] = s[n2][2] = s[n2][3] = 0.f;
                dp[n2][0] = dp[n2][1] = dp[n2][2] = dp[n2][3] = 0.f;
                mma_rows<D>(s[n2], ka, tQ, qc * 16 + n2 * 8, lane);
                mma_rows<D>(dp[n2], va, tDO, qc * 16 + n2 * 8, lane);
            }
            if (kTail && qc * 16 + 16 > lim) {   // queries beyond the sequence contribute nothing
#pragma unroll
                for (int n2 = 0; n2 < 2; ++n2) {
                    const int c = qc * 16 + n2 * 8 + 2 * t;
                    if (c >= lim) s[n2][0] = s[n2][2] = -INFINITY;
                    if (c + 1 >= lim) s[n2][1] = s[n2][3] = -INFINITY;
                }
            }
            uint32_t f0 = 0, f1 = 0;
            if (kDrop) {
                const uint32_t igk = ig + uint32_t(qc * 64 * 64) * kGolden;
                f0 = patch_flags(key, igk, p.keep_add);
                f1 = patch_flags(key, igk + 8u * kGolden, p.keep_add);
            }
            const int col = qc * 16 + 2 * t;
            bwd_chunk_math<kDrop, false, 0>(s[0], dp[0], p.scale_log2, *reinterpret_cast<const float2*>(Ls + col),
                                            *reinterpret_cast<const float2*>(Ds + col), f0, f1);
            bwd_chunk_math<kDrop, false, 1>(s[1], dp[1], p.scale_log2, *reinterpret_cast<const float2*>(Ls + col + 8),
                                            *reinterpret_cast<const float2*>(Ds + col + 8), f0, f1);
            // dp now holds P' = P * inv_keep; the dropped entries leave dV through the packed mask
            uint32_t a[4];
            a[0] = pack_bf16(dp[0][0], dp[0][1]);
            a[1] = pack_bf16(dp[0][2], dp[0][3]);
            a[2] = pack_bf16(dp[1][0], dp[1][1]);
            a[3] = pack_bf16(dp[1][2], dp[1][3]);
            if (kDrop) {
                a[0] &= pair_mask<0>(f0, f1);
                a[1] &= pair_mask<1>(f0, f1);
                a[2] &= pair_mask<2>(f0, f1);
                a[3] &= pair_mask<3>(f0, f1);
            }
            mma_cols<D>(dv, a, tDO, qc * 16, lane);
            a[0] = pack_bf16(s[0][0], s[0][1]);
            a[1] = pack_bf16(s[0][2], s[0][3]);
            a[2] = pack_bf16(s[1][0], s[1][1]);
            a[3] = pack_bf16(s[1][2], s[1][3]);
            mma_cols<D>(dk, a, tQ, qc * 16, lane);
        }
    }
}

// Role dK/dV: the warp owns 16 keys (K and V fragments in registers) and streams Q / dO / the two row statistics.
template <int D, bool kDrop>
__device__ __forceinline__ void attn_bwd_dkv(const AttnParams& p, unsigned char* smem, int seq, int head, int kt, int s0,
                                             int len) {
    constexpr int kTileBytes = kAttTile * D * 2;
    constexpr int kStatBytes = kAttTile * 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    constexpr int NS = att_stages<D>();
    const uint32_t sK = smem_u32(smem), sV = sK + kTileBytes, sQ = sV + kTileBytes, sDO = sQ + NS * kTileBytes,
                   sL = sDO + NS * kTileBytes, sD = sL + NS * kStatBytes;
    unsigned char* pL = smem + (2 + 2 * NS) * kTileBytes;
    unsigned char* pD = pL + NS * kStatBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(pD + NS * kStatBytes);
    const __nv_bfloat16* qs = p.q + (long long)s0 * p.in_stride + head * D;
    const __nv_bfloat16* ks = p.k + (long long)s0 * p.in_stride + head * D;
    const __nv_bfloat16* vs = p.v + (long long)s0 * p.in_stride + head * D;
    const __nv_bfloat16* dos = p.dout + (long long)s0 * (p.h * D) + head * D;
    const float* st_l = p.stat + (long long)head * p.T + s0;
    const float* st_d = st_l + (long long)p.h * p.T;
    const int nb = (len + kAttTile - 1) / kAttTile;
    const bool active = kt * kAttTile + warp * 16 < len;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) mbar_init(full + i, kAttThreads);
        fence_mbar_init();
    }
    __syncthreads();
    TileStream<D> tq, tdo;
    StatStream tl, td;
    tq.init(ks, p.in_stride, kt * kAttTile, tid);
    tq.load(sK, len);
    tq.init(vs, p.in_stride, kt * kAttTile, tid);
    tq.load(sV, len);
    tq.init(qs, p.in_stride, 0, tid);
    tdo.init(dos, (long long)p.h * D, 0, tid);
    tl.init(st_l, tid);
    td.init(st_d, tid);
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        if (i < nb) {
            tq.load(sQ + i * kTileBytes, len);
            tdo.load(sDO + i * kTileBytes, len);
            tl.load(sL + i * kStatBytes, len);
            td.load(sD + i * kStatBytes, len);
            cp_async_arrive(full + i);
        }
    }

    uint32_t ka[D / 16][4], va[D / 16][4];
    float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
        dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    uint32_t key = 0;
    if (kDrop) key = seq_key(p, seq, head);
    const uint32_t ig0 = patch_index(0, kt * 4 + warp, 2 * t, g) * kGolden;

    for (int qb = 0; qb < nb; ++qb) {
        const int st = qb % NS;
        if (active) {
            mbar_wait(full + st, (qb / NS) & 1);
            if (qb == 0) {
                load_a_frags<D>(sK, warp * 16, lane, ka);
                load_a_frags<D>(sV, warp * 16, lane, va);
            }
            const uint32_t tQ = sQ + st * kTileBytes, tDO = sDO + st * kTileBytes;
            const float* Ls = reinterpret_cast<const float*>(pL + st * kStatBytes);
            const float* Ds = reinterpret_cast<const float*>(pD + st * kStatBytes);
            const int lim = len - qb * kAttTile;
            const uint32_t ig = ig0 + uint32_t(qb * 4 * 64 * 64) * kGolden;
            if (SB200_ATT_BWD_SPLIT && lim >= kAttTile)
                bwd_dkv_block<D, kDrop, false>(p, ka, va, dk, dv, Ls, Ds, tQ, tDO, lim, key, ig, lane);
            else
                bwd_dkv_block<D, kDrop, true>(p, ka, va, dk, dv, Ls, Ds, tQ, tDO, lim, key, ig, lane);
        }
        if (qb + NS < nb) {
            __syncthreads();
            tq.load(sQ + st * kTileBytes, len);
            tdo.load(sDO + st * kTileBytes, len);
            tl.load(sL + st * kStatBytes, len);
            td.load(sD + st * kStatBytes, len);
            cp_async_arrive(full + st);
        }
    }
    if (!active) return;
    store_rows<D>(dk, p.scale, p.scale, smem, warp * 16, p.dk + (long long)s0 * p.d_stride + head * D, p.d_stride,
                  kt * kAttTile, len, lane);
    store_rows<D>(dv, 1.f, 1.f, smem + kTileBytes, warp * 16, p.dv + (long long)s0 * p.d_stride + head * D, p.d_stride,
                  kt * kAttTile, len, lane);
}

// kRole 0: both roles in one launch (blockIdx.x < ntile: dK/dV, the longer role, is scheduled first);
// 1 / 2: dK/dV only / dQ only (two launches: half the code per kernel).
template <int D, bool kDrop, int kRole>
__global__ void __launch_bounds__(kAttThreads, D == 32 ? SB200_ATT_BWD_CTAS32 : SB200_ATT_BWD_CTAS64) attn_bwd_kernel(const AttnParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int seq = blockIdx.z, head = blockIdx.y;
    const bool dkv = kRole == 1 || (kRole == 0 && int(blockIdx.x) < p.ntile);
    const int tile = (kRole == 0 && !dkv) ? blockIdx.x - p.ntile : blockIdx.x;
    const int s0 = __ldg(p.cu + seq);
    const int len = __ldg(p.cu + seq + 1) - s0;
    if (tile * kAttTile >= len) return;
    if (seq >= p.nseq_live) {
        if (dkv) {
            zero_rows<D>(p.dk + (long long)s0 * p.d_stride + head * D, p.d_stride, tile * kAttTile, len, threadIdx.x);
            zero_rows<D>(p.dv + (long long)s0 * p.d_stride + head * D, p.d_stride, tile * kAttTile, len, threadIdx.x);
        } else {
            zero_rows<D>(p.dq + (long long)s0 * p.d_stride + head * D, p.d_stride, tile * kAttTile, len, threadIdx.x);
        }
        return;
    }
    if (kRole != 2 && dkv) attn_bwd_dkv<D, kDrop>(p, smem, seq, head, tile, s0, len);
    if (kRole != 1 && !dkv) attn_bwd_dq<D, kDrop>(p, smem, seq, head, tile, s0, len);
}

// Test hook: mask[head, t, j] = 1 iff key j of token t's sequence is kept for query t (j < max_len).
__global__ void attn_mask_kernel(const AttnParams p, int max_len) {
    const int seq = blockIdx.z, head = blockIdx.y;
    const int s0 = p.cu[seq], len = p.cu[seq + 1] - s0;
    const uint32_t key = seq_key(p, seq, head);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < len * len; idx += gridDim.x * blockDim.x) {
        const int q = idx / len, k = idx % len;
        const int r = q & 15, c = k & 15;
        const uint32_t f = patch_flags(key, patch_index(q >> 4, k >> 4, r & 7, c & 7) * kGolden, p.keep_add);
        const int byte = (r >> 3) * 2 + (c >> 3);
        p.mask_out[((long long)head * p.T + s0 + q) * max_len + k] = (f >> (8 * byte + 7)) & 1u;
    }
}

int stages(int D) { return D == 32 ? att_stages<32>() : att_stages<64>(); }
size_t fwd_smem(int D) { return size_t(1 + 2 * stages(D)) * kAttTile * D * 2 + 64; }
size_t bwd_smem(int D) { return size_t(2 + 2 * stages(D)) * kAttTile * D * 2 + 2 * stages(D) * kAttTile * 4 + 64; }

int fill_dropout(AttnParams& p, float drop_p, const void* seed, int salt) {
    p.keep_thr = 256;
    p.keep_add = 0;
    p.inv_keep = 1.f;
    p.seed = nullptr;
    p.salt = uint32_t(salt);
    if (drop_p > 0.f) {
        SB200_REQUIRE(seed != nullptr, "attention: dropout needs a seed pointer");
        SB200_REQUIRE(drop_p <= 0.5f, "attention: drop_p must be <= 0.5");
        int thr = int((1.f - drop_p) * 256.f + 0.5f);
        thr = thr < 128 ? 128 : (thr > 256 ? 256 : thr);
        p.keep_thr = uint32_t(thr);
        p.keep_add = uint32_t(128 - (256 - thr)) * 0x01010101u;
        p.inv_keep = 256.f / float(thr);
        p.seed = static_cast<const unsigned long long*>(seed);
    }
    return SB200_OK;
}

template <typename K>
int opt_in_smem(K kernel, size_t bytes, int slot) {
    if (bytes > 48 * 1024 && slot >= 0 && !device_flag_test_and_set(slot))
        SB200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
    return SB200_OK;
}

int check_common(int T, int h, int d, int nseq, int max_len) {
    SB200_REQUIRE(d == 32 || d == 64, "attention: head_dim %d unsupported (32 or 64)", d);
    SB200_REQUIRE(T > 0 && h > 0 && h <= 65535 && nseq > 0 && nseq <= 65535, "attention: bad T / heads / nseq");
    SB200_REQUIRE(max_len > 0 && max_len <= 1024, "attention: max_len %d unsupported (1..1024)", max_len);
    return SB200_OK;
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" int sb200_attn_supported(int head_dim, int max_len) {
    return (head_dim == 32 || head_dim == 64) && max_len > 0 && max_len <= 1024;
}

extern "C" int sb200_attn_fwd(const void* q, const void* k, const void* v, size_t in_stride, const int* cu_seqlens,
                              int nseq, int nseq_live, int max_len, int T, int h, int d, float scale, float drop_p,
                              const void* drop_seed, int salt, void* out, float* lse, sb200_stream_t stream) {
    if (int rc = check_common(T, h, d, nseq, max_len)) return rc;
    SB200_REQUIRE(q && k && v && cu_seqlens && out && lse, "attn_fwd: null pointer");
    SB200_REQUIRE(nseq_live >= 0 && nseq_live <= nseq, "attn_fwd: nseq_live out of range");
    SB200_REQUIRE(in_stride % 8 == 0, "attn_fwd: row stride must be a multiple of 8 elements");
    AttnParams p{};
    p.q = static_cast<const __nv_bfloat16*>(q);
    p.k = static_cast<const __nv_bfloat16*>(k);
    p.v = static_cast<const __nv_bfloat16*>(v);
    p.in_stride = (long long)in_stride;
    p.out = static_cast<__nv_bfloat16*>(out);
    p.lse = lse;
    p.cu = cu_seqlens;
    p.T = T; p.h = h; p.nseq = nseq;
    p.nseq_live = nseq_live;
    p.ntile = (max_len + kAttTile - 1) / kAttTile;
    p.scale = scale;
    p.scale_log2 = scale * kLog2e;
    if (int rc = fill_dropout(p, drop_p, drop_seed, salt)) return rc;
    const dim3 grid(p.ntile, h, nseq);
    const size_t sm = fwd_smem(d);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool drop = p.keep_thr < 256;
#define SB200_ATT_FWD(DD, DR, SLOT)                                                            \
    do {                                                                                       \
        if (int rc = opt_in_smem(attn_fwd_kernel<DD, DR>, sm, SLOT)) return rc;                \
        attn_fwd_kernel<DD, DR><<<grid, kAttThreads, sm, s>>>(p);                              \
    } while (0)
    if (d == 32) { if (drop) SB200_ATT_FWD(32, true, -1); else SB200_ATT_FWD(32, false, -1); }   // 37 KB: no opt-in
    else { if (drop) SB200_ATT_FWD(64, true, 32); else SB200_ATT_FWD(64, false, 33); }          // 57 KB: opt-in slots 32 / 33
#undef SB200_ATT_FWD
    SB200_CHECK_LAUNCH("attn_fwd_kernel");
    return SB200_OK;
}

extern "C" int sb200_attn_bwd(const void* q, const void* k, const void* v, size_t in_stride, const void* out,
                              const void* dout, const float* lse, const int* cu_seqlens, int nseq, int nseq_live,
                              int max_len, int T,
                              int h, int d, float scale, float drop_p, const void* drop_seed, int salt, void* dq,
                              void* dk, void* dv, size_t d_stride, float* dsum, sb200_stream_t stream) {
    if (int rc = check_common(T, h, d, nseq, max_len)) return rc;
    SB200_REQUIRE(q && k && v && out && dout && lse && cu_seqlens && dq && dk && dv && dsum, "attn_bwd: null pointer");
    SB200_REQUIRE(nseq_live >= 0 && nseq_live <= nseq, "attn_bwd: nseq_live out of range");
    SB200_REQUIRE(in_stride % 8 == 0 && d_stride % 8 == 0, "attn_bwd: row strides must be multiples of 8 elements");
    AttnParams p{};
    p.q = static_cast<const __nv_bfloat16*>(q);
    p.k = static_cast<const __nv_bfloat16*>(k);
    p.v = static_cast<const __nv_bfloat16*>(v);
    p.in_stride = (long long)in_stride;
    p.lse = const_cast<float*>(lse);
    p.cu = cu_seqlens;
    p.T = T; p.h = h; p.nseq = nseq;
    p.nseq_live = nseq_live;
    p.ntile = (max_len + kAttTile - 1) / kAttTile;
    p.scale = scale;
    p.scale_log2 = scale * kLog2e;
    p.dout = static_cast<const __nv_bfloat16*>(dout);
    p.stat = dsum;
    p.dq = static_cast<__nv_bfloat16*>(dq);
    p.dk = static_cast<__nv_bfloat16*>(dk);
    p.dv = static_cast<__nv_bfloat16*>(dv);
    p.d_stride = (long long)d_stride;
    if (int rc = fill_dropout(p, drop_p, drop_seed, salt)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int prep_blocks = (T + 7) / 8;
    const __nv_bfloat16* o = static_cast<const __nv_bfloat16*>(out);
    if (d == 32)
        attn_prep_kernel<32><<<prep_blocks, 256, 0, s>>>(p.dout, o, lse, T, h, p.inv_keep, dsum);
    else
        attn_prep_kernel<64><<<prep_blocks, 256, 0, s>>>(p.dout, o, lse, T, h, p.inv_keep, dsum);
    SB200_CHECK_LAUNCH("attn_prep_kernel");
    const size_t sm = bwd_smem(d);
    const bool drop = p.keep_thr < 256;
#define SB200_ATT_BWD(DD, DR, ROLE, SLOT)                                                      \
    do {                                                                                       \
        if (int rc = opt_in_smem(attn_bwd_kernel<DD, DR, ROLE>, sm, SLOT)) return rc;          \
        attn_bwd_kernel<DD, DR, ROLE><<<grid, kAttThreads, sm, s>>>(p);                        \
    } while (0)
#if SB200_ATT_BWD_TWO_LAUNCHES
    const dim3 grid(p.ntile, h, nseq);
    if (d == 32) { if (drop) SB200_ATT_BWD(32, true, 1, -1); else SB200_ATT_BWD(32, false, 1, -1); }
    else { if (drop) SB200_ATT_BWD(64, true, 1, 14); else SB200_ATT_BWD(64, false, 1, 15); }
    SB200_CHECK_LAUNCH("attn_bwd_kernel");
    if (d == 32) { if (drop) SB200_ATT_BWD(32, true, 2, -1); else SB200_ATT_BWD(32, false, 2, -1); }
    else { if (drop) SB200_ATT_BWD(64, true, 2, 34); else SB200_ATT_BWD(64, false, 2, 35); }
#else
    const dim3 grid(2 * p.ntile, h, nseq);
    if (d == 32) { if (drop) SB200_ATT_BWD(32, true, 0, -1); else SB200_ATT_BWD(32, false, 0, -1); }
    else { if (drop) SB200_ATT_BWD(64, true, 0, 14); else SB200_ATT_BWD(64, false, 0, 15); }   // 67 KB: opt-in slots 14 / 15
#endif
#undef SB200_ATT_BWD
    SB200_CHECK_LAUNCH("attn_bwd_kernel");
    return SB200_OK;
}

extern "C" int sb200_attn_dropout_mask(const int* cu_seqlens, int nseq, int max_len, int T, int h, float drop_p,
                                       const void* drop_seed, int salt, unsigned char* mask,
                                       sb200_stream_t stream) {
    SB200_REQUIRE(cu_seqlens && mask && drop_seed, "attn_dropout_mask: null pointer");
    SB200_REQUIRE(drop_p > 0.f, "attn_dropout_mask: drop_p must be > 0");
    AttnParams p{};
    p.cu = cu_seqlens;
    p.T = T; p.h = h; p.nseq = nseq;
    if (int rc = fill_dropout(p, drop_p, drop_seed, salt)) return rc;
    p.mask_out = mask;
    attn_mask_kernel<<<dim3(8, h, nseq), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, max_len);
    SB200_CHECK_LAUNCH("attn_mask_kernel");
    return SB200_OK;
}
