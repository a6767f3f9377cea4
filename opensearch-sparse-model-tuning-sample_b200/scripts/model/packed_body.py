"""Padding-free execution of a HuggingFace BERT body (the third-party encoder body).

The reference pads every batch to its longest text and runs the whole encoder on the padding (``collator.py:34-41``,
``sparse_encoders.py:108``). Here the real tokens are packed into one ``[T_cap, H]`` matrix, every row-wise op
(Linear / GELU / LayerNorm / dropout, incl. the fused sm_100a kernels of ``fused_layers``) runs on packed rows only,
and attention runs per sequence on this repo's variable-length attention kernels (``ops.varlen_attention``,
csrc/attention.cu: head_dim 32 / 64, dropout on the probabilities; ``attention="flash"`` selects the ``flash_attn``
library kernel instead, for A/B runs and for head sizes the own kernels do not cover). Each block tail (dropout + residual add + LayerNorm + the bf16 cast for the next GEMM) is one sm_100a kernel
(``ops.add_layer_norm``). The packed result is scattered back to ``[B, L, H]`` for the fused sparse head.
The module re-uses the backbone's own sub-modules and parameters: nothing is copied, ``state_dict`` is unchanged.

Shapes are static (CUDA-graph friendly, no host synchronisation): ``T_cap = ceil(capacity * B * L)`` rows; tokens are
placed with device-side index arithmetic; the filler rows ``[T, T_cap)`` form dummy sequences so that every row stays
finite. ``capacity = 1.0`` can never overflow; a smaller capacity (chosen from the length distribution of the data)
also shrinks the GEMMs/elementwise work. An overflow (more real tokens than rows) cannot be handled on the fly without
a sync: it is recorded in ``overflow_count`` (device counter) and must be checked by the caller (the trainer does at
its logging steps, ``bench.py`` after the timed region).
"""
import math

import torch

from ... import ops

try:  # library kernel, optional (A/B runs, head sizes outside the own kernels)
    from flash_attn import flash_attn_varlen_qkvpacked_func
except Exception:  # pragma: no cover
    flash_attn_varlen_qkvpacked_func = None


class _Repad(torch.autograd.Function):
    """padded[b*L + l] = packed[dest[b, l]]; the backward is a gather through the inverse map (no atomics)."""

    @staticmethod
    def forward(ctx, packed, dest, src_of, row_valid):
        ctx.save_for_backward(src_of, row_valid)
        return packed.index_select(0, dest)

    @staticmethod
    def backward(ctx, g):
        src_of, row_valid = ctx.saved_tensors
        return g.index_select(0, src_of) * row_valid.to(g.dtype).unsqueeze(-1), None, None, None


class PackedBertBody:
    """Callable like ``backbone.bert(input_ids=..., attention_mask=...)[0]`` but padding-free inside. A plain object
    (not an nn.Module) so that the backbone's parameters are not registered twice."""

    def __init__(self, bert, capacity=1.0, attention="auto"):
        self.bert = bert
        self.capacity = float(capacity)
        cfg = bert.config
        self.num_heads = cfg.num_attention_heads
        self.head_dim = cfg.hidden_size // cfg.num_attention_heads
        self.attn_dropout = float(cfg.attention_probs_dropout_prob)
        if attention not in ("auto", "own", "flash"):
            raise ValueError(f"attention must be auto / own / flash, got {attention!r}")
        if attention == "flash" and flash_attn_varlen_qkvpacked_func is None:
            raise RuntimeError("attention='flash' needs the flash_attn package")
        if attention == "own" and (not ops.attn_supported(self.head_dim, 1) or self.attn_dropout > 0.5):
            raise RuntimeError(f"own attention kernels: head_dim {self.head_dim} / dropout {self.attn_dropout} unsupported")
        self.attention = attention
        self.overflow_count = None  # device int64 counter, created on first use
        self.step_overflow = None   # fp32 scalar device flag of the current step (the trainer resets and consumes it)

    @staticmethod
    def supported(backbone):
        bert = getattr(backbone, "bert", None)
        if bert is None:
            return False
        cfg = bert.config
        d = cfg.hidden_size // cfg.num_attention_heads
        own = ops.attn_supported(d, 1) and float(cfg.attention_probs_dropout_prob) <= 0.5
        if not own and flash_attn_varlen_qkvpacked_func is None:
            return False
        ok_cfg = getattr(cfg, "position_embedding_type", "absolute") == "absolute" and not getattr(cfg, "is_decoder", False)
        layer = bert.encoder.layer[0]
        return ok_cfg and hasattr(layer.attention, "self") and hasattr(layer.attention.self, "query") \
            and (cfg.hidden_size // cfg.num_attention_heads) in (32, 64, 96, 128, 192, 256) \
            and ops.layer_norm_supported(cfg.hidden_size)

    # ------------------------------------------------------------------------------------------------------------
    def _plan(self, attention_mask):
        """Device-side placement of the real tokens: no host sync, static shapes."""
        B, L = attention_mask.shape
        t_cap = int(math.ceil(self.capacity * B * L / 8.0) * 8)
        t_cap = max(8, min(t_cap, B * L))
        keep = attention_mask.reshape(-1) != 0
        rank_in_pack = torch.cumsum(keep.to(torch.int32), 0, dtype=torch.int32) - 1       # [B*L]
        total = rank_in_pack[-1] + 1                                                       # device scalar T
        fits = keep & (rank_in_pack < t_cap)
        dest = torch.where(fits, rank_in_pack, torch.full_like(rank_in_pack, t_cap)).long()  # dummy slot t_cap
        if self.overflow_count is None or self.overflow_count.device != keep.device:
            self.overflow_count = torch.zeros((), dtype=torch.int64, device=keep.device)
            self.step_overflow = torch.zeros((), dtype=torch.float32, device=keep.device)
        over = total > t_cap
        self.overflow_count += over.to(torch.int64)
        self.step_overflow.clamp_(min=over.to(torch.float32))
        # inverse map packed row -> padded position (filler rows point at 0 and are flagged invalid)
        flat = torch.arange(B * L, device=keep.device)
        src_of = torch.zeros(t_cap + 1, dtype=torch.long, device=keep.device).scatter_(0, dest, flat)[:t_cap]
        row_valid = torch.arange(t_cap, device=keep.device) < total
        # cumulative sequence lengths: B real sequences, then dummy sequences of <= L tokens covering the filler rows
        lens = attention_mask.ne(0).sum(1, dtype=torch.int32)
        cu_real = torch.cumsum(lens, 0, dtype=torch.int32).clamp_max(t_cap)
        n_dummy = (t_cap + L - 1) // L
        steps = torch.arange(1, n_dummy + 1, device=keep.device, dtype=torch.int32) * L
        cu_dummy = (total.clamp_max(t_cap) + steps).clamp_max(t_cap)   # (no host scalar -> tensor copies: graph capture)
        cu = torch.cat([torch.zeros(1, dtype=torch.int32, device=keep.device), cu_real, cu_dummy])
        return t_cap, dest, src_of, row_valid, cu

    def __call__(self, input_ids=None, attention_mask=None, token_type_ids=None, head_transform=None, **unused):
        """Returns (bf16 [T_cap, H] packed activations, plan). With `head_transform` (transformers
        BertPredictionHeadTransform) the MLM head transform (dense + activation + LayerNorm) is applied as well.
        Must run under bf16 autocast. Per block: one QKV GEMM on the concatenated projection weights, varlen
        attention on the packed qkv, and the fused dropout + residual + LayerNorm tail of `ops.add_layer_norm`
        that hands the next GEMM its bf16 operand (no separate cast / add / dropout kernels)."""
        B, L = input_ids.shape
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        t_cap, dest, src_of, row_valid, cu = self._plan(attention_mask)
        emb = self.bert.embeddings
        flat_ids = input_ids.reshape(-1)
        pos = torch.arange(L, device=input_ids.device).repeat(B)
        ids_p = torch.zeros(t_cap + 1, dtype=flat_ids.dtype, device=flat_ids.device).scatter_(0, dest, flat_ids)[:t_cap]
        pos_p = torch.zeros(t_cap + 1, dtype=pos.dtype, device=pos.device).scatter_(0, dest, pos)[:t_cap]
        if token_type_ids is None:
            typ_p = torch.zeros_like(ids_p)
        else:
            flat_t = token_type_ids.reshape(-1)
            typ_p = torch.zeros(t_cap + 1, dtype=flat_t.dtype, device=flat_t.device).scatter_(0, dest, flat_t)[:t_cap]
        tables = (emb.word_embeddings.weight, emb.position_embeddings.weight, emb.token_type_embeddings.weight)
        if all(t.dtype == torch.float32 for t in tables):
            x32 = ops.embed_sum(ids_p, pos_p, typ_p, *tables, padding_idx=emb.word_embeddings.padding_idx)
        else:
            x32 = emb.word_embeddings(ids_p) + emb.token_type_embeddings(typ_p) + emb.position_embeddings(pos_p)
        x32 = emb.dropout(emb.LayerNorm(x32)).float()
        x16 = x32.to(torch.bfloat16)

        h, d = self.num_heads, self.head_dim
        training = self.bert.training
        p_att = self.attn_dropout if training else 0.0
        scale = 1.0 / math.sqrt(d)
        layers = self.bert.encoder.layer
        own = self.attention == "own" or (self.attention == "auto" and ops.attn_supported(d, L)
                                          and self.attn_dropout <= 0.5)
        att_seed = None
        if own and p_att > 0.0:   # one seed per call from torch's generator (graph safe); the layer index is the salt
            att_seed = torch.randint(-2 ** 62, 2 ** 62, (1,), dtype=torch.int64, device=x16.device)
        for li, layer in enumerate(layers):
            att, sa = layer.attention, layer.attention.self
            qkv = ops.fused_qkv(x16, sa.query.weight, sa.key.weight, sa.value.weight, sa.query.bias, sa.key.bias,
                                sa.value.bias).view(t_cap, 3, h, d)
            if own:      # cu covers all t_cap rows (B real sequences, then the filler sequences, which are only zero-filled)
                ctx = ops.varlen_attention(qkv, cu, L, scale, p_att, training, seed=att_seed, salt=li,
                                           covers_all_rows=True, live_sequences=B)
            else:
                ctx = flash_attn_varlen_qkvpacked_func(qkv, cu, L, dropout_p=p_att, softmax_scale=scale, causal=False)
            y = att.output.dense(ctx.reshape(t_cap, h * d))
            ln = att.output.LayerNorm
            x32, x16 = ops.add_layer_norm(y, x32, ln.weight, ln.bias, ln.eps, p=att.output.dropout.p, training=training)
            y = layer.output.dense(self._dense_act(layer.intermediate.dense, layer.intermediate.intermediate_act_fn, x16))
            ln = layer.output.LayerNorm
            last = li == len(layers) - 1           # nothing reads the fp32 stream after the last block
            x32, x16 = ops.add_layer_norm(y, x32, ln.weight, ln.bias, ln.eps, p=layer.output.dropout.p,
                                          training=training, want_f32=not last)
        if head_transform is not None:
            y = self._dense_act(head_transform.dense, head_transform.transform_act_fn, x16)
            ln = head_transform.LayerNorm
            _, x16 = ops.add_layer_norm(y, None, ln.weight, ln.bias, ln.eps, want_f32=False)
        return x16, (dest.clamp_max(t_cap - 1), src_of, row_valid, (B, L), cu)

    @staticmethod
    def _dense_act(dense, act, x16):
        """act(dense(x)); the exact GELU of BERT runs on the fused sm_100a kernels (bias gradient included)."""
        exact_gelu = act is torch.nn.functional.gelu or (type(act).__name__ == "GELUActivation"
                                                         and getattr(act, "act", None) is torch.nn.functional.gelu)
        if exact_gelu and dense.out_features % 8 == 0 \
                and dense.out_features <= 4096:
            return ops.linear_gelu(x16, dense.weight, dense.bias)
        return act(dense(x16))

    @staticmethod
    def repad(packed, plan):
        dest, src_of, row_valid, (B, L) = plan[:4]
        return _Repad.apply(packed, dest, src_of, row_valid).view(B, L, packed.shape[-1])
