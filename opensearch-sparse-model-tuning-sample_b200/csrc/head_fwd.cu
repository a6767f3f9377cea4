// Fused sparse head forward for sm_100a.
//
// Reference semantics (scripts/model/sparse_encoders.py:108-114 + the MLM decoder Linear inside
// self.backbone): logits[b,l,v] = hidden[b,l,:].W[v,:] + bias[v]; values = max_l(logits * mask);
// rep = log1p(relu(values)) (log1p twice with use_l0).
//
// Mapping to the hardware
//   * one tcgen05.mma tile is D[128 vocab rows, N token columns] = W_tile[128, K] * hidden_tile[N, K]^T
//     (both operands K-major in shared memory, 128-byte swizzle, fed by TMA), fp32 accumulators in TMEM;
//   * vocab is the MMA M dimension, so TMEM lane i <-> vocab row i and TMEM column j <-> token j: the
//     max over the sequence is a per-thread running max over the columns a thread reads with
//     tcgen05.ld -- no shuffles, no shared memory, and the B x L x V logits never leave the SM;
//   * a token tile never straddles sequences: the hidden tensor map is 3-D [B, L, H] with box
//     {64, LC, S}; LC (chunk length, multiple of 16) and S (sequences per tile) are chosen on the host so
//     that N = S*LC <= 256. Out-of-range rows are zero-filled by TMA and masked out by the tile bitmap;
//   * bias is constant along l, so it is added after the max; log1p(relu(.)) is monotone and is applied
//     once per (b, v). Masked tokens contribute an exact 0 (the reference multiplies by the mask), which is
//     folded in as "max(x, 0) if the sequence has any masked slot";
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4..11 = two
//     epilogue warpgroups that alternate work units, double-buffered accumulators (2 x 256 TMEM columns);
//   * persistent CTAs (one per SM), static contiguous partition of the (sequence-group, vocab-tile) units.
#include <cstdlib>
#include <type_traits>
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "common.h"
#include "ptx.cuh"

namespace sb200 {

namespace {

constexpr int kBlockM = 128;     // vocab rows per tile (UMMA M)
constexpr int kBlockK = 64;      // bf16 per smem row = 128 B = swizzle span
constexpr int kUmmaK = 16;
constexpr int kMaxN = 256;       // token columns per tile (UMMA N upper bound)
constexpr int kNumEpiWG = 2;                 // epilogue warpgroups; every one of them works on every tile
constexpr int kNumEpiWarps = 4 * kNumEpiWG;  // 8 (16 were measured slower: register cap 96 + TMEM port contention)
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB per stage (this CTA's 128 vocab rows)
// B bytes per stage held by ONE CTA: the whole token tile (1-CTA) or half of it (CTA pair)
template <int kCG> struct StageCfg {
    static constexpr int kBBytes = kMaxN * kBlockK * 2 / kCG;   // 32 KB / 16 KB
    static constexpr int kStages = (kCG == 1) ? 4 : 6;          // 4 x 48 KB or 6 x 32 KB = 192 KB
    static constexpr size_t kSmemBytes = 1024 /*align slack*/ + size_t(kStages) * (kABytes + kBBytes) + 256 + 2 * kNumEpiWG * kBlockM * 8;
};
constexpr int kAccCols = 256;
constexpr int kTmemCols = 512;
constexpr int kFirstEpiWarp = 4;
constexpr int kThreads = (kFirstEpiWarp + kNumEpiWarps) * 32;  // 384
constexpr int kMaskWords = kMaxN / 32;                          // 8

struct HeadFwdParams {
    const float* bias;
    const uint32_t* tilemask;  // [n_groups][NC][8] validity bits per tile column
    const int4* seqinfo;       // [B] (number of masked slots, first masked slot, last real position + 1, packed start row)
    int packed;                // hidden is [T, H] (real tokens only, sequence b = rows seqinfo[b].w ...): 2-D tensor map
    float* rep;
    float* xmax;
    int32_t* argmax;
    int B, L, H, V;
    int LC, S, NC, N;          // chunk length, sequences per tile, chunks per sequence, S*LC
    int n_vtiles, n_groups, kblocks;   // n_vtiles counts tiles of 128 * kCG vocab rows
    int l0;
    int fp16;                  // operands are IEEE fp16 (else bf16)
    int n_peers;               // data-parallel all-gather fused into the epilogue: rep[b, v] is also stored into the
    float* peer_rep[7];        // gathered buffers of up to 7 other ranks (peer-mapped pointers to THIS rank's slot)
    int b_s_off;               // CTA pair, S >= 2: sequence offset of the second CTA's half of the token tile
};

// Packs the attention mask into per-tile column bitmaps and per-sequence padding info.
__global__ void head_prep_kernel(const void* __restrict__ mask, int elem_bytes, int B, int L, int LC, int S, int NC,
                                 int n_groups, uint32_t* __restrict__ tilemask, int4* __restrict__ seqinfo) {
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_words = n_groups * NC * kMaskWords;
    auto mask_at = [&](size_t i) -> bool {
        if (elem_bytes == 8) return reinterpret_cast<const int64_t*>(mask)[i] != 0;
        if (elem_bytes == 4) return reinterpret_cast<const int32_t*>(mask)[i] != 0;
        return reinterpret_cast<const uint8_t*>(mask)[i] != 0;
    };
    if (w < n_words) {
        const int word = w % kMaskWords;
        const int c = (w / kMaskWords) % NC;
        const int g = w / (kMaskWords * NC);
        const int n = word * 32 + lane;
        const int s = n / LC, j = n - s * LC;
        const int b = g * S + s, l = c * LC + j;
        bool valid = (s < S) && (b < B) && (l < L);
        if (valid) valid = mask_at(size_t(b) * L + l);
        const uint32_t bits = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) tilemask[w] = bits;
    } else if (w - n_words < B) {
        const int b = w - n_words;
        int count = 0, first = L, extent = 0;
        for (int l = lane; l < L; l += 32) {
            const bool v = mask_at(size_t(b) * L + l);
            count += v ? 1 : 0;
            if (!v && l < first) first = l;
            if (v) extent = l + 1;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            count += __shfl_xor_sync(0xffffffffu, count, o);
            first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
            extent = max(extent, __shfl_xor_sync(0xffffffffu, extent, o));
        }
        if (lane == 0) seqinfo[b] = make_int4(L - count, first, extent, 0);
    }
}

// Packed input (padding-free encoder body): sequence b is the run of rows [cu[b], cu[b+1]) of a [T, H] matrix, every
// row a real token. One sequence per tile (S == 1): validity bits and padding info follow from the lengths alone.
__global__ void head_prep_packed_kernel(const int32_t* __restrict__ cu, int B, int L, int LC, int NC,
                                        uint32_t* __restrict__ tilemask, int4* __restrict__ seqinfo) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_words = B * NC * kMaskWords;
    if (t < n_words) {
        const int word = t % kMaskWords;
        const int c = (t / kMaskWords) % NC;
        const int b = t / (kMaskWords * NC);
        const int len = min(max(__ldg(cu + b + 1) - __ldg(cu + b), 0), L);
        uint32_t bits = 0u;
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
            const int j = word * 32 + i;
            if (j < LC && c * LC + j < len) bits |= 1u << i;
        }
        tilemask[t] = bits;
    } else if (t - n_words < B) {
        const int b = t - n_words;
        const int start = __ldg(cu + b);
        const int len = min(max(__ldg(cu + b + 1) - start, 0), L);
        seqinfo[b] = make_int4(L - len, len, len, start);
    }
}

// Token columns a tile really needs (S == 1: one sequence per tile): padding beyond the last real token of the
// chunk is neither multiplied nor read back. Multiple of 16, at least 16.
__device__ __forceinline__ int tile_columns(const HeadFwdParams& p, int g, int c) {
    if (p.S != 1) return p.N;
    const int extent = __ldg(p.seqinfo + g).z - c * p.LC;
    const int n = (min(max(extent, 1), p.LC) + 15) & ~15;
    return n;
}

template <int kCG>
__global__ void __launch_bounds__(kThreads, 1)
head_fwd_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_h,
                const HeadFwdParams p) {
    constexpr int kStages = StageCfg<kCG>::kStages;
    constexpr int kBBytes = StageCfg<kCG>::kBBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + size_t(kStages) * kABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(kStages) * (kABytes + kBBytes));
    uint64_t* full_bar = bars;                   // [kStages]  TMA -> MMA
    uint64_t* empty_bar = bars + kStages;        // [kStages]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * kStages;    // [2]        MMA -> epilogue
    uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]     epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    float2* comb = reinterpret_cast<float2*>(bars + 2 * kStages + 6);  // [2][kNumEpiWG][128] partial (max, argmax)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // CTA pair (kCG == 2): rank 0 is the leader (issues the MMAs, owns the full/tempty barriers that are waited on)
    const uint32_t rank = (kCG == 2) ? cluster_ctarank() : 0u;
    const int cluster_id = blockIdx.x / kCG;
    const int n_clusters = gridDim.x / kCG;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_h);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], (blockDim.x / 32 - kFirstEpiWarp) * kCG);  // one arrive per epilogue warp (of both CTAs)
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        if (kCG == 2) {
            tmem_alloc_pair(tmem_slot, kTmemCols);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_slot, kTmemCols);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (kCG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long total_units = (long long)p.n_vtiles * p.n_groups;
    const int u_begin = int((long long)cluster_id * total_units / n_clusters);
    const int u_end = int((long long)(cluster_id + 1) * total_units / n_clusters);
    // bytes landing per stage over the whole CTA group (credited to the leader's barrier)
    const uint32_t tx_bytes = uint32_t(kCG * kABytes + p.N * kBlockK * 2);

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer
            uint32_t it = 0;
            for (int u = u_begin; u < u_end; ++u) {
                const int g = u / p.n_vtiles, vt = u - g * p.n_vtiles;
                for (int c = 0; c < p.NC; ++c) {
                    // CTA pair, one sequence per tile: the second CTA's half starts where the first one's ends
                    const int l_off = (p.S == 1) ? tile_columns(p, g, c) / 2 : 0;
                    const int row0 = p.packed ? __ldg(p.seqinfo + g).w + c * p.LC : 0;   // packed: first row of the chunk
                    for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
                        const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        if (kCG == 1) {
                            mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
                            tma_load_2d(smem_a + s * kABytes, &tmap_w, &full_bar[s], kb * kBlockK, vt * kBlockM);
                            if (p.packed) tma_load_2d(smem_b + s * kBBytes, &tmap_h, &full_bar[s], kb * kBlockK, row0);
                            else tma_load_3d(smem_b + s * kBBytes, &tmap_h, &full_bar[s], kb * kBlockK, c * p.LC, g * p.S);
                        } else {
                            // each CTA loads its 128 vocab rows and its half of the token tile
                            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
                            tma_load_2d_pair(smem_a + s * kABytes, &tmap_w, &full_bar[s], kb * kBlockK,
                                             (vt * 2 + int(rank)) * kBlockM);
                            if (p.packed)
                                tma_load_2d_pair(smem_b + s * kBBytes, &tmap_h, &full_bar[s], kb * kBlockK,
                                                 row0 + int(rank) * l_off);
                            else
                                tma_load_3d_pair(smem_b + s * kBBytes, &tmap_h, &full_bar[s], kb * kBlockK,
                                                 c * p.LC + int(rank) * l_off, g * p.S + int(rank) * p.b_s_off);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // ------------------------------------------------ MMA issuer (single thread of the leader CTA)
            uint32_t it = 0, ci = 0;
            // the column count of the next tile is fetched one tile ahead (global load off the issue path)
            int n_next = (u_begin < u_end) ? tile_columns(p, u_begin / p.n_vtiles, 0) : 16;
            for (int u = u_begin; u < u_end; ++u) {
                for (int c = 0; c < p.NC; ++c, ++ci) {
                    const uint32_t idesc = umma_idesc_f16(kBlockM * kCG, n_next, p.fp16 != 0);
                    {
                        int un = u, cn = c + 1;
                        if (cn == p.NC) { cn = 0; ++un; }
                        if (un < u_end) n_next = tile_columns(p, un / p.n_vtiles, cn);
                    }
                    const uint32_t as = ci & 1, aph = (ci >> 1) & 1;
                    mbar_wait(&tempty_bar[as], aph ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + as * kAccCols;
                    for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
                        const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                        mbar_wait(&full_bar[s], ph);
                        tc_fence_after();
                        const uint64_t da = umma_desc_sw128(smem_u32(smem_a + s * kABytes));
                        const uint64_t db = umma_desc_sw128(smem_u32(smem_b + s * kBBytes));
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                            // advance 32 bytes (= 16 bf16) inside the 128-byte swizzle row
                            if (kCG == 1)
                                umma_bf16(tmem_d, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                            else
                                umma_bf16_pair(tmem_d, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc,
                                               (kb | k) != 0 ? 1u : 0u);
                        }
                        // frees the smem stage (in both CTAs of a pair) when these MMAs retire
                        if (kCG == 1) umma_commit(&empty_bar[s]); else umma_commit_pair(&empty_bar[s], 3);
                    }
                    // accumulator ready for the epilogue (of both CTAs)
                    if (kCG == 1) umma_commit(&tfull_bar[as]); else umma_commit_pair(&tfull_bar[as], 3);
                }
            }
        }
    } else if (warp >= kFirstEpiWarp) {
        // ---------------------------------------------------- epilogue: all warpgroups share EVERY accumulator tile
        // (the epilogue of a tile has to fit inside the MMA time of the next one: two accumulator stages).
        const int ew = warp - kFirstEpiWarp;
        const int wg = ew >> 2;
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = uint32_t(quarter * 32) << 16;
        uint32_t ci = 0;
        float m = -CUDART_INF_F;
        int idx = 0;

        const bool want_arg = p.argmax != nullptr;
        // value-only variant (inference: no arg-max wanted): 8 three-input max instructions per 16 columns
        auto block16_value = [&](const uint32_t (&r)[16], uint32_t bits, auto masked) {
            constexpr bool kMasked = decltype(masked)::value;
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                v[i] = __uint_as_float(r[i]);
                if (kMasked && !((bits >> i) & 1u)) v[i] = -CUDART_INF_F;
            }
            const float a0 = max3f(v[0], v[1], v[2]), a1 = max3f(v[3], v[4], v[5]), a2 = max3f(v[6], v[7], v[8]);
            const float a3 = max3f(v[9], v[10], v[11]), a4 = max3f(v[12], v[13], v[14]);
            m = max3f(max3f(a0, a1, a2), max3f(a3, a4, v[15]), m);
        };
        // 16-column block: tree arg-max (depth 4) merged into the running (m, idx); strict '>' keeps the lowest
        // position on ties. kMasked: columns whose bit is clear are excluded.
        auto block16 = [&](const uint32_t (&r)[16], uint32_t bits, int lbase, auto masked) {
            constexpr bool kMasked = decltype(masked)::value;
            float v[16];
            int ix[8];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                v[i] = __uint_as_float(r[i]);
                if (kMasked && !((bits >> i) & 1u)) v[i] = -CUDART_INF_F;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool pr = v[2 * i + 1] > v[2 * i];
                ix[i] = pr ? 2 * i + 1 : 2 * i;
                v[i] = fmaxf(v[2 * i], v[2 * i + 1]);
            }
#pragma unroll
            for (int w = 4; w >= 1; w >>= 1) {
#pragma unroll
                for (int i = 0; i < w; ++i) {
                    const bool pr = v[2 * i + 1] > v[2 * i];
                    ix[i] = pr ? ix[2 * i + 1] : ix[2 * i];
                    v[i] = fmaxf(v[2 * i], v[2 * i + 1]);
                }
            }
            if (v[0] > m) {
                m = v[0];
                idx = lbase + ix[0];
            }
        };
        // columns [n_begin, n_end) of the accumulator (multiples of 16) hold tokens l_begin.. of one sequence
        auto scan_columns = [&](uint32_t tmem_acc, const uint32_t* tm, int n_begin, int n_end, int l_begin) {
            for (int n0 = n_begin; n0 < n_end; n0 += 32) {
                const int n1 = n0 + 16;
                const uint32_t lo = (__ldg(tm + (n0 >> 5)) >> (n0 & 31)) & 0xffffu;
                const uint32_t hi = (n1 < n_end) ? ((__ldg(tm + (n1 >> 5)) >> (n1 & 31)) & 0xffffu) : 0u;
                uint32_t ra[16], rb[16];
                if (lo != 0) tmem_ld16(tmem_acc + n0, ra);
                if (hi != 0) tmem_ld16(tmem_acc + n1, rb);
                if ((lo | hi) != 0) tmem_ld_wait();
                const int l0 = l_begin + (n0 - n_begin);
                if (!want_arg) {
                    if (lo == 0xffffu) block16_value(ra, lo, std::false_type{});
                    else if (lo != 0) block16_value(ra, lo, std::true_type{});
                    if (hi == 0xffffu) block16_value(rb, hi, std::false_type{});
                    else if (hi != 0) block16_value(rb, hi, std::true_type{});
                    continue;
                }
                if (lo == 0xffffu) block16(ra, lo, l0, std::false_type{});
                else if (lo != 0) block16(ra, lo, l0, std::true_type{});
                if (hi == 0xffffu) block16(rb, hi, l0 + 16, std::false_type{});
                else if (hi != 0) block16(rb, hi, l0 + 16, std::true_type{});
            }
        };
        auto finalize = [&](int b, int v, float bias_v) {
            const int4 si = __ldg(p.seqinfo + b);
            float x = m + bias_v;
            if (si.x > 0) {
                if (x < 0.f || (x == 0.f && si.y < idx)) idx = si.y;
                x = fmaxf(x, 0.f);
            }
            const size_t o = size_t(b) * p.V + v;
            if (p.xmax != nullptr) p.xmax[o] = x;
            if (p.argmax != nullptr) p.argmax[o] = idx;
            float r1 = log1pf(fmaxf(x, 0.f));
            if (p.l0) r1 = log1pf(r1);
            p.rep[o] = r1;
            // fused all-gather: the same value goes straight into every peer's gathered buffer over NVLink (a warp writes
            // 32 consecutive floats = one 128-byte line per peer), overlapped with the MMAs of the following tiles
#pragma unroll 1
            for (int k = 0; k < p.n_peers; ++k) p.peer_rep[k][o] = r1;
        };

        // Work split of one tile over the n_wg warpgroups:
        //   S >= n_wg sequences per tile: whole sequences round-robin, nothing to merge;
        //   fewer: each sequence is cut into n_wg/S column ranges, one per warpgroup; the partial (max, argmax) pairs
        //          go through shared memory and one warpgroup of the group (rotating per unit) merges and writes.
        const int n_wg = int(blockDim.x / 32 - kFirstEpiWarp) / 4;   // = kNumEpiWG
        const bool split = p.S < n_wg && n_wg % p.S == 0;         // a sequence is shared by several warpgroups
        const int gsz = split ? n_wg / p.S : 1;
        const int my_s = wg / gsz, part = wg - my_s * gsz;
        for (int u = u_begin; u < u_end; ++u) {
            const int g = u / p.n_vtiles, vt = u - g * p.n_vtiles;
            const int v = (vt * kCG + int(rank)) * kBlockM + row;
            const bool v_ok = v < p.V;
            const float bias_v = (v_ok && p.bias != nullptr) ? __ldg(p.bias + v) : 0.f;
            if (split) {
                m = -CUDART_INF_F;
                idx = 0;
            }
            for (int c = 0; c < p.NC; ++c, ++ci) {
                const uint32_t as = ci & 1, aph = (ci >> 1) & 1;
                const uint32_t* tm = p.tilemask + (size_t(g) * p.NC + c) * kMaskWords;
                const int n_seq = (p.S == 1) ? tile_columns(p, g, c) : p.LC;  // columns worth scanning per sequence
                mbar_wait(&tfull_bar[as], aph);
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + lane_base + as * kAccCols;
                if (!split) {
                    for (int s = wg; s < p.S; s += n_wg) {
                        m = -CUDART_INF_F;
                        idx = 0;
                        scan_columns(tmem_acc, tm, s * p.LC, (s + 1) * p.LC, 0);
                        const int b = g * p.S + s;
                        if (b < p.B && v_ok) finalize(b, v, bias_v);
                    }
                } else {
                    const int nb = n_seq >> 4;  // 16-column blocks of this sequence in this chunk
                    const int b0 = part * nb / gsz, b1 = (part + 1) * nb / gsz;
                    const int base = my_s * p.LC;
                    scan_columns(tmem_acc, tm, base + 16 * b0, base + 16 * b1, c * p.LC + 16 * b0);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (kCG == 1) mbar_arrive(&tempty_bar[as]); else mbar_arrive_cluster(&tempty_bar[as], 0);
                }
            }
            if (split) {
                // merge the partial results of the group: larger value wins, equal values keep the lower position
                const int ui = u - u_begin;
                float2* buf = comb + (ui & 1) * (kNumEpiWG * kBlockM);
                const int merger = my_s * gsz + ui % gsz;
                if (wg != merger) buf[wg * kBlockM + row] = make_float2(m, __int_as_float(idx));
                asm volatile("bar.sync 1, %0;" ::"r"(n_wg * 128) : "memory");
                if (wg == merger) {
                    for (int k = my_s * gsz; k < (my_s + 1) * gsz; ++k) {
                        if (k == wg) continue;
                        const float2 o = buf[k * kBlockM + row];
                        const int oi = __float_as_int(o.y);
                        if (o.x > m || (o.x == m && oi < idx)) {
                            m = o.x;
                            idx = oi;
                        }
                    }
                    const int b = g * p.S + my_s;
                    if (b < p.B && v_ok) finalize(b, v, bias_v);
                }
            }
        }
    }

    tc_fence_before();
    if (kCG == 2) cluster_sync_all(); else __syncthreads();  // the peer may still signal barriers in this CTA's smem
    if (warp == 2) {
        tc_fence_after();
        if (kCG == 2) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess) return nullptr;
        if (qres != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

struct HeadTiling {
    int LC, S, NC, N, n_groups;
};

// pair != 0: the token tile must split into two equal halves along whole sequences (S even) or inside one (S == 1)
HeadTiling head_tiling(int B, int L, int pair = 0) {
    HeadTiling t;
    if (L <= kMaxN) {
        t.NC = 1;
        t.LC = int(align_up(size_t(L), 16));
        t.S = kMaxN / t.LC;
        if (t.S > B) t.S = B;
        if (t.S < 1) t.S = 1;
        if (pair && t.S > 1 && (t.S & 1)) t.S -= 1;
    } else {
        t.NC = (L + kMaxN - 1) / kMaxN;
        t.LC = int(align_up(size_t((L + t.NC - 1) / t.NC), 16));
        t.S = 1;
    }
    t.N = t.S * t.LC;
    t.n_groups = (B + t.S - 1) / t.S;
    return t;
}

}  // namespace

}  // namespace sb200

using namespace sb200;

static size_t head_ws_bytes(const HeadTiling& t, int B) {
    return align_up(size_t(t.n_groups) * t.NC * kMaskWords * sizeof(uint32_t), 256) + align_up(size_t(B) * sizeof(int4), 256);
}

// 1 = single-CTA tiles (128 x N), 2 = CTA pairs (256 x N, cta_group::2). SB200_HEAD_CTA_GROUP overrides (A/B testing).
static int head_cta_group() {
    static int cached = [] {
        const char* e = getenv("SB200_HEAD_CTA_GROUP");
        if (e != nullptr && (e[0] == '1' || e[0] == '2')) return e[0] - '0';
        return 2;
    }();
    return cached;
}

extern "C" size_t sb200_head_fwd_workspace_bytes(int B, int L) {
    if (B <= 0 || L <= 0) return 0;
    const size_t a = head_ws_bytes(head_tiling(B, L, 0), B), b = head_ws_bytes(head_tiling(B, L, 1), B);
    return a > b ? a : b;
}

// cu_seqlens == nullptr: hidden is the padded [B, L, H] tensor and `mask` the attention mask; otherwise hidden is the
// packed [T, H] matrix of real tokens and sequence b its rows [cu_seqlens[b], cu_seqlens[b+1]) (one sequence per tile).
static int head_fwd_impl(const void* hidden, const void* W, const float* bias, const void* mask, int mask_elem_bytes,
                         const int32_t* cu_seqlens, int T, int B, int L, int H, int V, int flags, float* rep, float* xmax,
                         int32_t* argmax, float* const* peer_rep, int n_peers, void* workspace, size_t workspace_bytes,
                         sb200_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const bool packed = cu_seqlens != nullptr;
    SB200_REQUIRE(hidden && W && (mask || packed) && rep, "head_fwd: null pointer");
    SB200_REQUIRE(n_peers >= 0 && n_peers <= 7 && (n_peers == 0 || peer_rep != nullptr), "head_fwd: bad peer list (%d)",
                  n_peers);
    SB200_REQUIRE(B >= 1 && V >= 1 && L >= 1 && L <= 4096, "head_fwd: bad shape B=%d L=%d V=%d", B, L, V);
    SB200_REQUIRE(H >= 8 && H % 8 == 0, "head_fwd: H=%d must be a positive multiple of 8", H);
    SB200_REQUIRE(packed || mask_elem_bytes == 1 || mask_elem_bytes == 4 || mask_elem_bytes == 8,
                  "head_fwd: mask_elem_bytes=%d", mask_elem_bytes);
    SB200_REQUIRE(!packed || (L > 128 && T >= 1), "head_fwd_packed: needs max_len > 128 (one sequence per tile), T >= 1");
    SB200_REQUIRE((reinterpret_cast<uintptr_t>(hidden) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
                  "head_fwd: hidden/W must be 16-byte aligned");
    const size_t need = sb200_head_fwd_workspace_bytes(B, L);
    if (workspace == nullptr || workspace_bytes < need)
        return fail(SB200_ERR_WORKSPACE, "head_fwd: workspace %zu < %zu", workspace_bytes, need);

    const int sms = num_sms();
    const int cg = (head_cta_group() == 2 && sms >= 2) ? 2 : 1;
    const HeadTiling t = head_tiling(B, L, cg == 2);
    uint32_t* tilemask = static_cast<uint32_t*>(workspace);
    int4* seqinfo = reinterpret_cast<int4*>(static_cast<uint8_t*>(workspace) +
                                            align_up(size_t(t.n_groups) * t.NC * kMaskWords * sizeof(uint32_t), 256));

    EncodeTiledFn encode = get_encode_fn();
    if (encode == nullptr) return fail(SB200_ERR_CUDA, "head_fwd: cuTensorMapEncodeTiled unavailable");

    // the B box is what ONE CTA loads per stage: the whole token tile, or (CTA pair) half of it
    // (one sequence per tile: half of the chunk per CTA, the second half's start is chosen per tile in the kernel)
    int box_l = t.LC, box_s = t.S, b_s_off = 0;
    if (cg == 2) {
        if (t.S >= 2) { box_s = t.S / 2; b_s_off = t.S / 2; }
        else          { box_l = t.LC / 2; }
    }
    CUtensorMap tmap_w, tmap_h;
    const CUtensorMapDataType operand_type =
        (flags & SB200_HEAD_FP16) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    {
        cuuint64_t dims[2] = {cuuint64_t(H), cuuint64_t(V)};
        cuuint64_t strides[1] = {cuuint64_t(H) * 2};
        cuuint32_t box[2] = {kBlockK, kBlockM};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmap_w, operand_type, 2, const_cast<void*>(W), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SB200_ERR_CUDA, "head_fwd: tensor map (W) encode failed: %d", int(r));
    }
    if (packed) {
        cuuint64_t dims[2] = {cuuint64_t(H), cuuint64_t(T)};
        cuuint64_t strides[1] = {cuuint64_t(H) * 2};
        cuuint32_t box[2] = {kBlockK, cuuint32_t(box_l)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmap_h, operand_type, 2, const_cast<void*>(hidden), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SB200_ERR_CUDA, "head_fwd: tensor map (packed hidden) encode failed: %d", int(r));
    } else {
        cuuint64_t dims[3] = {cuuint64_t(H), cuuint64_t(L), cuuint64_t(B)};
        cuuint64_t strides[2] = {cuuint64_t(H) * 2, cuuint64_t(L) * cuuint64_t(H) * 2};
        cuuint32_t box[3] = {kBlockK, cuuint32_t(box_l), cuuint32_t(box_s)};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&tmap_h, operand_type, 3, const_cast<void*>(hidden), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SB200_ERR_CUDA, "head_fwd: tensor map (hidden) encode failed: %d", int(r));
    }

    if (packed) {
        if (t.S != 1) return fail(SB200_ERR_ARG, "head_fwd_packed: tiling holds %d sequences per tile", t.S);
        const int n = B * t.NC * kMaskWords + B;
        head_prep_packed_kernel<<<(n + 255) / 256, 256, 0, stream>>>(cu_seqlens, B, L, t.LC, t.NC, tilemask, seqinfo);
        SB200_CHECK_LAUNCH("head_prep_packed_kernel");
    } else {
        const int n_warps = t.n_groups * t.NC * kMaskWords + B;
        const int threads = 256;
        const int blocks = (n_warps * 32 + threads - 1) / threads;
        head_prep_kernel<<<blocks, threads, 0, stream>>>(mask, mask_elem_bytes, B, L, t.LC, t.S, t.NC, t.n_groups,
                                                         tilemask, seqinfo);
        SB200_CHECK_LAUNCH("head_prep_kernel");
    }

    HeadFwdParams p;
    p.bias = bias;
    p.tilemask = tilemask;
    p.seqinfo = seqinfo;
    p.packed = packed ? 1 : 0;
    p.rep = rep;
    p.xmax = xmax;
    p.argmax = argmax;
    p.B = B; p.L = L; p.H = H; p.V = V;
    p.LC = t.LC; p.S = t.S; p.NC = t.NC; p.N = t.N;
    p.n_vtiles = (V + kBlockM * cg - 1) / (kBlockM * cg);
    p.n_groups = t.n_groups;
    p.kblocks = (H + kBlockK - 1) / kBlockK;
    p.l0 = (flags & SB200_HEAD_L0) ? 1 : 0;
    p.fp16 = (flags & SB200_HEAD_FP16) ? 1 : 0;
    p.n_peers = n_peers;
    for (int k = 0; k < 7; ++k) p.peer_rep[k] = k < n_peers ? peer_rep[k] : nullptr;
    p.b_s_off = b_s_off;

    const long long total_units = (long long)p.n_vtiles * p.n_groups;
    if (cg == 1) {
        // per-device attribute, set once per device (never while a stream capture may be in progress later on)
        if (!device_flag_test_and_set(0))
            SB200_CUDA(cudaFuncSetAttribute(head_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            int(StageCfg<1>::kSmemBytes)));
        int grid = sms;
        if (grid > total_units) grid = int(total_units);
        head_fwd_kernel<1><<<grid, kThreads, StageCfg<1>::kSmemBytes, stream>>>(tmap_w, tmap_h, p);
    } else {
        if (!device_flag_test_and_set(5))
            SB200_CUDA(cudaFuncSetAttribute(head_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            int(StageCfg<2>::kSmemBytes)));
        int clusters = sms / 2;
        if (clusters > total_units) clusters = int(total_units);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(unsigned(clusters * 2));
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = StageCfg<2>::kSmemBytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        SB200_CUDA(cudaLaunchKernelEx(&cfg, head_fwd_kernel<2>, tmap_w, tmap_h, p));
    }
    SB200_CHECK_LAUNCH("head_fwd_kernel");
    return SB200_OK;
}

extern "C" int sb200_head_fwd(const void* hidden, const void* W, const float* bias, const void* mask,
                              int mask_elem_bytes, int B, int L, int H, int V, int flags, float* rep, float* xmax,
                              int32_t* argmax, float* const* peer_rep, int n_peers, void* workspace,
                              size_t workspace_bytes, sb200_stream_t stream) {
    return head_fwd_impl(hidden, W, bias, mask, mask_elem_bytes, nullptr, 0, B, L, H, V, flags, rep, xmax, argmax, peer_rep,
                         n_peers, workspace, workspace_bytes, stream);
}

extern "C" int sb200_head_fwd_packed(const void* hidden, const void* W, const float* bias, const int32_t* cu_seqlens, int T,
                                     int B, int max_len, int H, int V, int flags, float* rep, float* xmax, int32_t* argmax,
                                     float* const* peer_rep, int n_peers, void* workspace, size_t workspace_bytes,
                                     sb200_stream_t stream) {
    if (cu_seqlens == nullptr) return fail(SB200_ERR_ARG, "head_fwd_packed: cu_seqlens is null");
    return head_fwd_impl(hidden, W, bias, nullptr, 0, cu_seqlens, T, B, max_len, H, V, flags, rep, xmax, argmax, peer_rep,
                         n_peers, workspace, workspace_bytes, stream);
}
