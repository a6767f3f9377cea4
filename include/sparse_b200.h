/*
 * sparse_b200.h -- C ABI of the B200 (sm_100a) neural-sparse encoding hot path.
 *
 * Every entry point replaces one PyTorch-op cluster of the reference
 * (zhichao-aws/opensearch-sparse-model-tuning-sample); the reference location each one stands in
 * for is cited per function as file:line relative to the reference root.
 *
 * Conventions (all functions):
 *   - plain C types only: raw DEVICE pointers, sizes, and a cudaStream_t passed as void*;
 *   - the caller owns every buffer (inputs, outputs, workspace); nothing is allocated inside;
 *   - launches are asynchronous on `stream`; no host synchronisation happens inside;
 *   - re-entrant, no global mutable state apart from a thread-local last-error string;
 *   - return value: SB200_OK (0) or an SB200_ERR_* code; sb200_last_error() describes the failure.
 *   - matrices are dense row-major; "bf16" is the raw 16-bit bfloat16 pattern.
 */
#ifndef SPARSE_B200_H_
#define SPARSE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_ABI_VERSION 3

enum {
    SB200_OK = 0,
    SB200_ERR_ARG = 1,         /* bad shape / null pointer / unsupported size */
    SB200_ERR_CUDA = 2,        /* a CUDA runtime/driver call failed */
    SB200_ERR_WORKSPACE = 3    /* workspace too small */
};

/* flags for the sparse head */
enum {
    SB200_HEAD_L0 = 1,         /* second log1p ("use_l0", sparse_encoders.py:113-114) */
    SB200_HEAD_FP16 = 2        /* hidden and W are IEEE fp16 instead of bf16 (the reference configs train with fp16: true);
                                  products are exact and accumulated in fp32 either way */
};

/* loss modes for sb200_score_loss_* (scripts/train/loss.py:110 LOSS_CLS_MAP) */
enum {
    SB200_LOSS_INFONCE = 0,    /* loss.py:80-107 */
    SB200_LOSS_KLDIV = 1,      /* loss.py:18-43  */
    SB200_LOSS_MARGINMSE = 2   /* loss.py:46-77  */
};

typedef void* sb200_stream_t;  /* cudaStream_t */

int sb200_abi_version(void);
const char* sb200_last_error(void);
/* Number of kernels this library has launched in the calling process (for bench.py's gpu_launches). */
unsigned long long sb200_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * (1) Fused sparse head forward.   Replaces sparse_encoders.py:108-114 (MLM decoder GEMM inside
 *     self.backbone + `output * mask` + torch.max(dim=1) + log1p(relu) [+ log1p]) and
 *     bi_encoder_wrapper.py:29-33.
 *
 *   hidden  bf16 [B, L, H]   output of the MLM head transform (input of the vocab decoder); fp16 with SB200_HEAD_FP16
 *   W       bf16 [V, H]      decoder weight (tied word embeddings); fp16 with SB200_HEAD_FP16
 *   bias    f32  [V]         decoder bias (may be NULL)
 *   mask    [B, L] attention mask, element size mask_elem_bytes in {1, 4, 8}; non-zero = real token
 *   rep     f32  [B, V]  out: log1p(relu(max_l(logit*mask)))  (log1p applied twice with SB200_HEAD_L0)
 *   xmax    f32  [B, V]  out (nullable): the pooled pre-activation value max_l(logit*mask)
 *   argmax  i32  [B, V]  out (nullable): sequence position of that maximum (lowest l on ties, the
 *                        first masked position when the maximum is the 0 of a masked slot)
 *   peer_rep  host array of n_peers (0..7) device pointers, or NULL: the fused all-gather of gather_rep
 *             (scripts/utils.py:16-23). Entry k is THIS rank's [B, V] slot inside rank k's gathered buffer
 *             (peer-mapped through sb200_peer_import); the epilogue stores every finished rep[b, v] there as
 *             well, over NVLink, while the following tiles are being multiplied. Publish with
 *             sb200_peer_signal, consume after sb200_peer_wait.
 *   H % 8 == 0, 1 <= L <= 4096, V >= 1.  Logits never touch HBM.
 * ------------------------------------------------------------------------------------------- */
size_t sb200_head_fwd_workspace_bytes(int B, int L);
int sb200_head_fwd(const void* hidden, const void* W, const float* bias, const void* mask, int mask_elem_bytes,
                   int B, int L, int H, int V, int flags, float* rep, float* xmax, int32_t* argmax,
                   float* const* peer_rep, int n_peers, void* workspace, size_t workspace_bytes,
                   sb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (2) Sparse head backward.   Replaces the autograd backward of sparse_encoders.py:108-114
 *     (dense dlogits + two dense GEMMs) with a gather/scatter over the B*V winning positions.
 *
 *   d_rep   f32 [B, V]   gradient w.r.t. rep
 *   xmax, argmax         as saved by sb200_head_fwd
 *   d_hidden f32 [B, L, H] out (fully written)       dW f32 [V, H] out (fully written)
 *   dbias   f32 [V] out (nullable)
 * ------------------------------------------------------------------------------------------- */
size_t sb200_head_bwd_workspace_bytes(int B, int L, int H, int V);
int sb200_head_bwd(const float* d_rep, const float* xmax, const int32_t* argmax, const void* hidden, const void* W,
                   int B, int L, int H, int V, int flags, float* d_hidden, float* dW, float* dbias,
                   void* workspace, size_t workspace_bytes, sb200_stream_t stream);

/* (1p)/(2p) The same head on PACKED hidden states (padding-free encoder body): hidden is [T, H], every row a real token,
 * sequence b = rows [cu_seqlens[b], cu_seqlens[b+1]) (device i32 [B+1], lengths <= max_len). No padded [B, L, H] copy
 * is made in either direction: the forward reads the runs through a 2-D tensor map (one sequence per tile, so
 * max_len > 128), the backward gathers / scatters packed rows and returns d_hidden as [T, H] f32. A sequence shorter
 * than max_len counts as having masked slots (their logit * mask = 0 takes part in the max, as in the reference);
 * argmax is the token's rank inside its sequence. sb200_head_packed_supported(H, max_len) tells whether the pair of
 * calls is available (H in {64, 128, 256, 384, 512, 768}, 128 < max_len <= 4096). Workspaces as for the padded calls
 * with L = max_len. */
int sb200_head_packed_supported(int H, int max_len);
int sb200_head_fwd_packed(const void* hidden, const void* W, const float* bias, const int32_t* cu_seqlens, int T, int B,
                          int max_len, int H, int V, int flags, float* rep, float* xmax, int32_t* argmax,
                          float* const* peer_rep, int n_peers, void* workspace, size_t workspace_bytes,
                          sb200_stream_t stream);
int sb200_head_bwd_packed(const float* d_rep, const float* xmax, const int32_t* argmax, const void* hidden, const void* W,
                          const int32_t* cu_seqlens, int T, int B, int max_len, int H, int V, int flags, float* d_hidden,
                          float* dW, float* dbias, void* workspace, size_t workspace_bytes, sb200_stream_t stream);

/* (2b) Row pruning, in place: rep[b,v] *= (rep[b,v] > ratio * max_v rep[b,:]).  sparse_encoders.py:115-119 */
int sb200_prune_rows(float* rep, int B, int V, float ratio, sb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (3) Inf-free IDF query encoding.   Replaces sparse_encoders.py:121-127.
 *   ids      i64 or i32 [Nq, Lq] token ids, ids_elem_bytes = 8 or 4 (attention mask is ignored, as in the reference)
 *   idf      f32 [V]        idf_vector parameter
 *   special  i32 [n_special] token ids forced to zero
 *   q        f32 [Nq, V]    out: relu(idf[v]) where v occurs in row and is not special, else 0
 * Bit-exact with the reference.  Ids outside [0, V) are an error in the reference (index error);
 * here they are ignored and counted in *bad_ids (device i32, nullable).
 * ------------------------------------------------------------------------------------------- */
int sb200_idf_query(const void* ids, int ids_elem_bytes, const float* idf, const int32_t* special, int n_special, int Nq,
                    int Lq, int V, float* q, int32_t* bad_ids, sb200_stream_t stream);
/* d_idf[v] = sum_b d_q[b,v] * [q[b,v] > 0]   (gradient of the above w.r.t. idf; idf_requires_grad) */
int sb200_idf_query_bwd(const float* d_q, const float* q, int Nq, int V, float* d_idf, sb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (4) FLOPS / L0-thresholded FLOPS regulariser.   Replaces trainer.py:61-73 (flops_value).
 *   rep      f32 [N*G, V] viewed as [N, G, V]
 *   threshold < 0: plain FLOPS  sum_{g,v} (mean_n |rep|)^2
 *   threshold >= 0: only rows with nnz(row) > threshold contribute to the mean (mean still over N)
 *   colsum   f32 [G, V]  out: sum_n rowmask*|rep|  (saved for backward)
 *   rowmask  f32 [N*G]   out: 1/0 row mask; written only with a threshold or stats (may be NULL otherwise; the
 *                        backward treats a NULL rowmask as all ones)
 *   value    f32 [1]     out
 *   stats    f32 [4]     out (nullable): total nnz, sum of positive entries, max entry, unused
 *   workspace: sb200_flops_workspace_bytes(N, G, V). Plain FLOPS is ONE launch (column owners with batched row loads,
 *   deterministic last-block sum of the value); the thresholded variant adds the row pass.
 * Backward: d_rep[n,g,v] (+)= gscale * 2*colsum[g,v]/N^2 * sign(rep) * rowmask   (gscale read on device)
 * ------------------------------------------------------------------------------------------- */
size_t sb200_flops_workspace_bytes(int N, int G, int V);
int sb200_flops_fwd(const float* rep, int N, int G, int V, float threshold, float* colsum, float* rowmask,
                    float* value, float* stats, void* workspace, size_t workspace_bytes, sb200_stream_t stream);
int sb200_flops_bwd(const float* rep, const float* colsum, const float* rowmask, const float* gscale, int N, int G,
                    int V, int row_begin, int row_end, int accumulate, float* d_rep, sb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (5) Query x doc score matrix.   Replaces the matmul/bmm of loss.py:30-37, 62-68, 92-101 and
 *     bi_encoder_wrapper.py:124-131.   S[i,j] = sum_v q[i,v] * d[j,v], fp32 accumulate.
 *   in_batch != 0: S is [Nq, Nd] (all pairs).  in_batch == 0: S is [Nq, G], G = Nd/Nq, query i
 *   against its own docs i*G .. i*G+G-1.
 * ------------------------------------------------------------------------------------------- */
size_t sb200_scores_workspace_bytes(int Nq, int Nd, int V, int in_batch);
/* in_batch with a workspace: the query rows are thresholded (!= 0) into ordered (column, value) lists (one block per
 * query, no atomics) and every document row is streamed through shared memory once (ring of bulk-async-copied row parts,
 * any V) and gathered by all query lists: one HBM pass over d, no atomics, deterministic. Lists of up to 1024 entries
 * per query are supported (short ones live in registers, longer ones are streamed from L2); beyond that a device-side
 * flag routes the work to the dense fp32 tile kernel (no host sync). Without a workspace only the dense kernel runs.
 * q_nnz_bound > 0 promises that no query row has more non-zeros than that (0 = unknown): the dense fallback kernels are
 * then not even enqueued. in_batch == 0: one cluster of CTAs per query (DSMEM reduction). */
int sb200_scores_fwd(const float* q, const float* d, int Nq, int Nd, int V, int in_batch, int q_nnz_bound, float* S,
                     void* workspace, size_t workspace_bytes, sb200_stream_t stream);
/* d_q[i,:] = gs * sum_j dS[i,j] d[j,:] (rows q_begin..q_end) ; d_d[j,:] = gs * sum_i dS[i,j] q[i,:] (rows
 * d_begin..d_end); gs = *gscale (device scalar, nullable = 1): the upstream gradient of a scalar loss, so that no
 * separate dS * g pass is needed. Either output may be NULL. accumulate != 0 adds into the outputs. Outputs are
 * full-size [Nq,V] / [Nd,V]. fwd_workspace (nullable): the workspace the forward call filled for the same q; enables
 * the sparse-query path. */
int sb200_scores_bwd(const float* dS, const float* gscale, const float* q, const float* d, int Nq, int Nd, int V,
                     int in_batch, int q_begin, int q_end, int d_begin, int d_end, int accumulate, float* d_q,
                     float* d_d, const void* fwd_workspace, size_t workspace_bytes, sb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (6) Ranking losses on a score matrix.   Replaces loss.py:33-42 (KLDiv), 64-76 (MarginMSE),
 *     94-106 (InfoNCE incl. the boolean-mask negative gather and the one-hot CE).
 *   S        f32 [Nq, C]   C = Nd (in_batch) or G
 *   teacher  f32 [Nq, C]   (kldiv / marginmse; NULL for infonce)
 *   G        docs per query (positive first); in_batch infonce drops other queries' positives
 *   loss     f32 [1] out ;  dS f32 [Nq, C] out (nullable): d loss / d S
 *   workspace: sb200_rank_loss_workspace_bytes(Nq) (row losses; the last block adds them in row order: deterministic)
 * ------------------------------------------------------------------------------------------- */
size_t sb200_rank_loss_workspace_bytes(int Nq);
int sb200_rank_loss(int mode, const float* S, const float* teacher, int Nq, int C, int G, int in_batch,
                    float temperature, float* loss, float* dS, void* workspace, size_t workspace_bytes,
                    sb200_stream_t stream);

/* (5)+(6) in one call: S = q.d^T, loss = ranking loss(S [, teacher]) and dS = d loss / d S, i.e. what
 * LOSS_CLS_MAP[name](...).__call__(q_rep, d_rep, inputs) computes (loss.py:25-43, 57-77, 86-107), Nd == Nq * G.
 * in_batch: 2 launches (query lists + one persistent cooperative kernel: document rows -> grid barrier -> loss rows
 * spread over the CTAs -> last block sums). own docs: 1 launch (cluster per query). q_nnz_bound > 0 promises that no
 * query row has more non-zeros than that (inf-free queries: the token count); 0 = unknown, the dense fallback kernels
 * are then enqueued behind a device-side flag. workspace: sb200_scores_workspace_bytes(Nq, Nd, V, in_batch); pass the
 * same workspace to sb200_scores_bwd. */
int sb200_score_loss_fwd(int mode, const float* q, const float* d, const float* teacher, int Nq, int Nd, int V, int G,
                         int in_batch, float temperature, int q_nnz_bound, float* S, float* loss, float* dS,
                         void* workspace, size_t workspace_bytes, sb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (7) "next" rows: encode output path (sparse_encoders.py:137-150, 178-179) and the teacher
 *     ensemble normalisation (bi_encoder_wrapper.py:133-138).
 * ------------------------------------------------------------------------------------------- */
/* CSR compaction of the non-zero entries of columns >= first_col (first_col = 1 reproduces the post-processor's
 * column-0 sentinel being dropped), columns ascending inside a row; also df_count[v] += [rep[b,v] > 0] for EVERY
 * column v (i64, nullable; sparse_encoders.py:178-179).
 * row_ptr i32 [B+1]; cols i32 / vals f32 sized capacity; row_ptr[B] = total nnz (entries past capacity are dropped). */
size_t sb200_compact_workspace_bytes(int B, int V);
int sb200_compact_rows(const float* rep, int B, int V, int first_col, int32_t* row_ptr, int32_t* cols, float* vals,
                       int capacity, int64_t* df_count, void* workspace, size_t workspace_bytes, sb200_stream_t stream);
/* acc[i,:] (+)= scale * (S[i,:] - min_i) / (max_i - min_i + 1e-6) */
int sb200_minmax_accumulate(const float* S, int Nq, int C, float scale, int accumulate, float* acc,
                            sb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (8) "next" row, rank 4 (encoder body): fused LayerNorm forward / backward replacing
 *     torch.nn.functional.layer_norm inside the third-party backbone (transformers BertLayer /
 *     BertEmbeddings / BertPredictionHeadTransform), whose backward was the largest single item of the
 *     training step once the head is fused.
 *   x, y, dy, dx  [R, H] bf16 (elem_bytes 2) or fp32 (elem_bytes 4);  gamma, beta, dgamma, dbeta f32 [H]
 *   mean, rstd    f32 [R] saved by the forward for the backward.  H % 128 == 0, H/128 in {1,2,3,4,6,8}.
 * ------------------------------------------------------------------------------------------- */
int sb200_layer_norm_supported(int H);
int sb200_layer_norm_fwd(const void* x, int elem_bytes, const float* gamma, const float* beta, int R, int H, float eps,
                         void* y, float* mean, float* rstd, sb200_stream_t stream);
size_t sb200_layer_norm_bwd_workspace_bytes(int R, int H);
int sb200_layer_norm_bwd(const void* x, const void* dy, int elem_bytes, const float* gamma, const float* mean,
                         const float* rstd, int R, int H, void* dx, float* dgamma, float* dbeta, void* workspace,
                         size_t workspace_bytes, sb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Residual tail of a transformer block: out = LayerNorm(dropout(y) + resid), delivered as fp32 and as bf16.
 *   replaces transformers BertSelfOutput.forward / BertOutput.forward (dropout, residual add, LayerNorm) as run by
 *   the backbone call at scripts/model/sparse_encoders.py:108 under bf16 autocast, plus the fp32->bf16 casts autocast
 *   inserts in front of the following Linear layers, and the matching autograd chain.
 *   y        bf16 [R, H]   branch output (before dropout)
 *   resid    f32  [R, H]   residual stream, or NULL (plain LayerNorm of dropout(y))
 *   drop_seed  device pointer to one 64-bit seed, or NULL / drop_p == 0 for no dropout. The keep mask is a pure
 *            function of (seed, element index) (Philox4x32-10) and is regenerated by the backward call: pass the
 *            same seed and drop_p.
 *   out_f32 / out_bf16   either may be NULL (not both); out_bf16 = round-to-nearest of out_f32
 *   mean, rstd  f32 [R]    saved statistics
 * Backward: g_f32 / g_bf16 = gradients w.r.t. the two outputs (either NULL, summed in fp32); writes d_y (bf16),
 * d_resid (f32, may be NULL), dgamma / dbeta (f32 [H], deterministic order); workspace as sb200_layer_norm_bwd. */
int sb200_add_layer_norm_fwd(const void* y, const float* resid, const float* gamma, const float* beta, int R, int H,
                             float eps, const void* drop_seed, float drop_p, float* out_f32, void* out_bf16,
                             float* mean, float* rstd, sb200_stream_t stream);
int sb200_add_layer_norm_bwd(const void* y, const float* resid, const float* g_f32, const void* g_bf16,
                             const float* gamma, const float* mean, const float* rstd, int R, int H,
                             const void* drop_seed, float drop_p, void* d_y, float* d_resid, float* dgamma,
                             float* dbeta, void* workspace, size_t workspace_bytes, sb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Exact GELU y = x * Phi(x) on bf16 and its backward fused with the bias gradient of the Linear in front of it:
 *   replaces transformers BertIntermediate.intermediate_act_fn / BertPredictionHeadTransform.transform_act_fn
 *   (torch.nn.functional.gelu) inside the backbone call at scripts/model/sparse_encoders.py:108, its autograd
 *   backward, and the bias-gradient reduction of the preceding torch.nn.Linear.
 *   x, y, dy, dx  bf16 [R, N] (gelu_fwd: any n % 8 == 0 elements);  colsum f32 [N] = sum_r dx[r, :] or NULL.
 *   N % 8 == 0, N <= 4096 for the backward. Phi via erfc (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7). */
int sb200_gelu_fwd(const void* x, size_t n, void* y, sb200_stream_t stream);
size_t sb200_gelu_bwd_workspace_bytes(int R, int N);
int sb200_gelu_bwd(const void* x, const void* dy, int R, int N, void* dx, float* colsum, void* workspace,
                   size_t workspace_bytes, sb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Input embeddings of the body on (packed) token rows and the gradient scatter into the three tables:
 *   replaces transformers BertEmbeddings.forward (word + token-type + position lookups and their sum) inside the
 *   backbone call at scripts/model/sparse_encoders.py:108 and the three embedding_dense_backward pipelines.
 *   ids, pos, typ  int64 [n] (clamped to the table sizes);  W [nW, H], P [nP, H], T [nT, H] fp32;  out, g fp32 [n, H]
 *   out[t] = (W[ids[t]] + T[typ[t]]) + P[pos[t]].   Backward ADDS into dW / dP / dT (caller zeroes them); tokens with
 *   ids == pad_idx (torch.nn.Embedding padding_idx, -1 = none) leave dW untouched. H % 4 == 0. */
int sb200_embed_sum_fwd(const int64_t* ids, const int64_t* pos, const int64_t* typ, const float* W, const float* P,
                        const float* T, int n, int H, int nW, int nP, int nT, float* out, sb200_stream_t stream);
int sb200_embed_sum_bwd(const int64_t* ids, const int64_t* pos, const int64_t* typ, const float* g, int n, int H,
                        int nW, int nP, int nT, int pad_idx, float* dW, float* dP, float* dT, sb200_stream_t stream);

/* Column sums of a row-major [R, N] matrix (bf16 or fp32): out[c] = sum_r dy[r,c]. The bias gradient of the body's
 * Linear layers (replaces the torch reduce kernel behind addmm's backward). N % 8 == 0, N <= 4096. */
int sb200_colsum_supported(int N);
size_t sb200_colsum_workspace_bytes(int R, int N);
int sb200_colsum(const void* dy, int elem_bytes, int R, int N, float* out, void* workspace, size_t workspace_bytes,
                 sb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Variable-length multi-head self-attention of the body on packed token rows, forward and backward:
 *   replaces transformers BertSelfAttention.forward (scores = Q K^T / sqrt(d), softmax, attention-probs dropout,
 *   probs V) inside the backbone call at scripts/model/sparse_encoders.py:108 and its autograd chain; the padding the
 *   reference multiplies (collator.py:34-41 pads to the longest text) is never touched: sequence i is token rows
 *   [cu_seqlens[i], cu_seqlens[i + 1]) and attends only to itself (non-causal, no further mask).
 *   q, k, v     bf16, element [t, head, :] at ptr + t * in_stride + head * d (a fused [T, 3, h, d] projection output:
 *               q = base, k = base + h * d, v = base + 2 * h * d, in_stride = 3 * h * d); in_stride % 8 == 0
 *   cu_seqlens  int32 [nseq + 1] on the device, non-decreasing, cu_seqlens[nseq] <= T; empty sequences allowed
 *   nseq_live   sequences [nseq_live, nseq) are filler (the unused capacity of a packed batch): their rows of out /
 *               dq / dk / dv are zero-filled (lse = 0) and no attention is computed for them
 *   max_len     host upper bound of every sequence length (<= 1024); d = 32 or 64
 *   out         bf16 [T, h * d];  lse f32 [h, T] = log sum_j exp(scale * q . k_j)
 *   drop_p / drop_seed / salt: dropout on the probabilities. The keep mask is a pure function of (the 64-bit value at
 *               drop_seed, salt, sequence, head, query, key); keep probability is round(256 * (1 - drop_p)) / 256 and
 *               the kept entries are rescaled by its inverse. Pass the same three values to the backward call.
 *               drop_p == 0: no dropout (drop_seed may be NULL).
 * Backward: dout bf16 [T, h * d]; writes dq / dk / dv (bf16, element [t, head, :] at ptr + t * d_stride + head * d,
 * every row of every sequence exactly once; rows outside all sequences are left untouched); dsum f32 [2, h, T] is
 * scratch (per-row constants of the pass). Deterministic (no atomics). drop_p <= 0.5.
 * sb200_attn_dropout_mask (test hook): mask u8 [h, T, max_len], mask[head, t, j] = 1 iff query t keeps key j of its
 * own sequence. */
int sb200_attn_supported(int head_dim, int max_len);
int sb200_attn_fwd(const void* q, const void* k, const void* v, size_t in_stride, const int* cu_seqlens, int nseq,
                   int nseq_live, int max_len, int T, int h, int d, float scale, float drop_p, const void* drop_seed, int salt,
                   void* out, float* lse, sb200_stream_t stream);
int sb200_attn_bwd(const void* q, const void* k, const void* v, size_t in_stride, const void* out, const void* dout,
                   const float* lse, const int* cu_seqlens, int nseq, int nseq_live, int max_len, int T, int h, int d,
                   float scale,
                   float drop_p, const void* drop_seed, int salt, void* dq, void* dk, void* dv,
                   size_t d_stride, float* dsum, sb200_stream_t stream);
int sb200_attn_dropout_mask(const int* cu_seqlens, int nseq, int max_len, int T, int h, float drop_p,
                            const void* drop_seed, int salt, unsigned char* mask, sb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Symmetric peer memory over NVLink (CUDA IPC), one process per GPU: the exchange steps of data-parallel training
 * without NCCL -- gather_rep (scripts/utils.py:16-23 = accelerate's gather), the teacher gather
 * (bi_encoder_wrapper.py:130) and the score gather (trainer.py:101-104).
 *   alloc / export / import / close / free: every rank allocates the same buffer (zero-filled), exports a 64-byte IPC
 *   handle, imports its peers' handles (peer access is enabled on import).
 *   allgather: copies `bytes` (multiple of 16) from src into slot `rank` (offset dst_byte_offset + rank * bytes) of every
 *   rank's buffer; dst_ptrs is a HOST array of `world` (<= 16) device pointers, entry p = rank p's buffer as mapped here.
 *   signal: epoch = ++(*send_epoch) (device counter), then flags_of_rank_p[rank] = epoch for every p (release, system
 *   scope); flag_ptrs like dst_ptrs. wait: epoch = ++(*wait_epoch), spins until all `world` local flags reached it.
 *   All three are kernels on the caller's stream (CUDA-graph capturable, no host synchronisation). A buffer may be
 *   rewritten once every rank has consumed it; in the training step the gradient all-reduce between two steps orders
 *   that. */
int sb200_peer_alloc(size_t bytes, void** ptr);
int sb200_peer_free(void* ptr);
int sb200_peer_export(const void* ptr, void* handle64);
int sb200_peer_import(const void* handle64, void** ptr);
int sb200_peer_close(void* ptr);
int sb200_peer_allgather(const void* src, size_t bytes, int rank, int world, void* const* dst_ptrs,
                         size_t dst_byte_offset, sb200_stream_t stream);
int sb200_peer_signal(uint32_t* send_epoch, int rank, int world, void* const* flag_ptrs, size_t flag_byte_offset,
                      sb200_stream_t stream);
int sb200_peer_wait(uint32_t* wait_epoch, const uint32_t* local_flags, int world, sb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPARSE_B200_H_ */
