"""CPU oracle of the neural-sparse encoding hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain fp32 torch-on-CPU arithmetic, what the reference
(zhichao-aws/opensearch-sparse-model-tuning-sample) computes on the path this repository accelerates. It is the
checker for the CUDA kernels: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it. The product package never does (and fails loudly without its CUDA
library).

Parity pinning: the reference ships no tests or golden vectors, so the oracle is pinned against
outputs of the reference's own modules, generated in the build container by ``oracle/make_golden.py`` (which imports
``/root/reference``) and committed under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.

Every function names the reference lines it follows (paths relative to the reference root).
"""

import torch

NEG_INF = float("-inf")


# ----------------------------------------------------------------------------------------------- sparse head
def pooled_logits(logits, attention_mask):
    """max over the sequence of mask-multiplied logits -> (values [B,V], argmax [B,V]).

    scripts/model/sparse_encoders.py:109-111 (and bi_encoder_wrapper.py:30-32). Masked slots contribute an exact 0,
    not -inf; ties resolve to the lowest position, as torch.max does on CPU.
    """
    weighted = logits.float() * attention_mask.to(logits.device).unsqueeze(-1).float()
    B, L, V = weighted.shape
    best = weighted[:, 0, :].clone()
    where = torch.zeros(B, V, dtype=torch.long)
    for l in range(1, L):
        row = weighted[:, l, :]
        better = row > best
        best = torch.where(better, row, best)
        where = torch.where(better, torch.full_like(where, l), where)
    return best, where


def activation(values, use_l0):
    """log1p(relu(x)), applied a second time for the L0 variant.  sparse_encoders.py:112-114."""
    out = torch.log1p(values.clamp_min(0.0))
    return torch.log1p(out) if use_l0 else out


def prune(rep, prune_ratio):
    """Keep entries strictly above prune_ratio * row max.  sparse_encoders.py:115-119."""
    if prune_ratio is None:
        return rep
    limit = rep.amax(dim=-1, keepdim=True) * prune_ratio
    return rep * (rep > limit).to(rep.dtype)


def decoder_logits(hidden, weight, bias):
    """The MLM decoder Linear(H, V) that ends self.backbone(**kw)[0] (transformers BertLMPredictionHead.decoder,
    called at sparse_encoders.py:108). fp32."""
    out = hidden.float() @ weight.float().t()
    return out if bias is None else out + bias.float()


def sparse_head(hidden, weight, bias, attention_mask, use_l0=False, prune_ratio=None):
    """sparse_encoders.py:107-119 from the decoder input onwards -> (rep, values, argmax)."""
    values, where = pooled_logits(decoder_logits(hidden, weight, bias), attention_mask)
    return prune(activation(values, use_l0), prune_ratio), values, where


def sparse_head_grads(hidden, weight, bias, attention_mask, d_rep, use_l0=False):
    """Autograd reference for the head: grads of sum(rep * d_rep) w.r.t. (hidden, weight, bias), fp32."""
    h = hidden.float().clone().requires_grad_(True)
    w = weight.float().clone().requires_grad_(True)
    b = None if bias is None else bias.float().clone().requires_grad_(True)
    logits = h @ w.t()
    if b is not None:
        logits = logits + b
    masked = logits * attention_mask.unsqueeze(-1).float()
    values = masked.max(dim=1).values
    rep = torch.log1p(torch.relu(values))
    if use_l0:
        rep = torch.log1p(rep)
    (rep * d_rep).sum().backward()
    return h.grad, w.grad, (None if b is None else b.grad)


def teacher_sparse_head(hidden, weight, bias, attention_mask, special_token_ids):
    """BiSparseModel.forward: single log(1+relu), special-token columns zeroed.  bi_encoder_wrapper.py:28-35."""
    values, _ = pooled_logits(decoder_logits(hidden, weight, bias), attention_mask)
    out = torch.log(1.0 + values.clamp_min(0.0))
    out[:, list(special_token_ids)] = 0.0
    return out


# ----------------------------------------------------------------------------------------------- inf-free query
def idf_query(input_ids, idf_vector, special_token_ids):
    """Bag of token ids weighted by relu(idf); attention mask ignored.  sparse_encoders.py:121-127."""
    n, V = input_ids.shape[0], idf_vector.shape[0]
    present = torch.zeros(n, V, dtype=torch.float32)
    present.scatter_(1, input_ids.long(), 1.0)
    present[:, list(special_token_ids)] = 0.0
    return present * idf_vector.float().clamp_min(0.0)


def build_idf_vector(idf_items, token_to_id, vocab_size):
    """Default weight 1.0, overwritten per token.  sparse_encoders.py:86-91."""
    vec = torch.ones(vocab_size, dtype=torch.float32)
    for token, weight in idf_items:
        vec[token_to_id(token)] = weight
    return vec


# ----------------------------------------------------------------------------------------------- regularisers
def flops_value(rep, group_num=1, flops_threshold=None):
    """FLOPS regulariser, optionally restricted to rows longer than flops_threshold.  trainer.py:61-73."""
    V = rep.shape[-1]
    grouped = rep.float().abs().reshape(-1, group_num, V)
    if flops_threshold is not None:
        length = (grouped != 0).sum(dim=2)
        keep = (length > flops_threshold).to(grouped.dtype).unsqueeze(2)
        grouped = grouped * keep
    return (grouped.mean(dim=0) ** 2).sum()


def get_lambda(lambda_value, lambda_T, global_step):
    """Quadratic warm-up of the regulariser weight.  trainer.py:75-79."""
    if global_step >= lambda_T:
        return lambda_value
    return lambda_value * ((global_step + 1) / lambda_T) ** 2


# ----------------------------------------------------------------------------------------------- ranking losses
def _own_doc_scores(q_rep, d_rep):
    """bmm branch: query i against its own G docs -> [Nq, G].  loss.py:28-35, 60-66, 100-101."""
    nq, V = q_rep.shape
    docs = d_rep.reshape(nq, -1, V)
    return torch.einsum("igv,iv->ig", docs.float(), q_rep.float())


def student_scores(q_rep, d_rep, use_in_batch_negatives):
    if use_in_batch_negatives:
        return q_rep.float() @ d_rep.float().t()
    return _own_doc_scores(q_rep, d_rep)


def infonce_loss(q_rep, d_rep, use_in_batch_negatives=False):
    """Cross-entropy with the positive in column 0; in-batch negatives are all hard negatives of all queries,
    other queries' positives are excluded.  loss.py:86-107."""
    nq = q_rep.shape[0]
    G = d_rep.shape[0] // nq
    full = q_rep.float() @ d_rep.float().t()
    pos_cols = torch.arange(nq) * G
    pos = full[torch.arange(nq), pos_cols]
    is_neg = torch.ones(d_rep.shape[0], dtype=torch.bool)
    is_neg[pos_cols] = False
    if use_in_batch_negatives:
        neg = full[:, is_neg]
    else:
        own = _own_doc_scores(q_rep, d_rep)
        neg = own[:, 1:]
    logits = torch.cat([pos.unsqueeze(1), neg], dim=1)
    return (torch.logsumexp(logits, dim=1) - logits[:, 0]).mean()


def kldiv_loss(q_rep, d_rep, teacher_scores, use_in_batch_negatives=False, temperature=1.0):
    """KL(teacher || student) on temperature-scaled softmaxes, summed over docs, mean over queries.  loss.py:25-43."""
    s = student_scores(q_rep, d_rep, use_in_batch_negatives) / temperature
    t = teacher_scores.float() / temperature
    log_ps = s - torch.logsumexp(s, dim=1, keepdim=True)
    log_pt = t - torch.logsumexp(t, dim=1, keepdim=True)
    pt = log_pt.exp()
    return (pt * (log_pt - log_ps)).sum(dim=1).mean()


def marginmse_loss(q_rep, d_rep, teacher_scores, use_in_batch_negatives=False, temperature=1.0):
    """MSE between student and teacher margins (column 0 minus every other column).  loss.py:57-77."""
    s = student_scores(q_rep, d_rep, use_in_batch_negatives) / temperature
    t = teacher_scores.float() / temperature
    ms = s[:, :1] - s[:, 1:]
    mt = t[:, :1] - t[:, 1:]
    return ((ms - mt) ** 2).mean()


LOSSES = {"infonce": infonce_loss, "kldiv": kldiv_loss, "marginmse": marginmse_loss}


def ranking_loss(name, q_rep, d_rep, teacher_scores=None, use_in_batch_negatives=False, temperature=1.0, weight=1.0):
    """SparseTrainingLoss.get_loss = weight * loss.  loss.py:14-15, 110."""
    if name == "infonce":
        return weight * infonce_loss(q_rep, d_rep, use_in_batch_negatives)
    return weight * LOSSES[name](q_rep, d_rep, teacher_scores, use_in_batch_negatives, temperature)


# ----------------------------------------------------------------------------------------------- cross-rank gather
def gather_rep(local_reps, rank):
    """What rank `rank` sees after gather_rep: rank-major concat of every rank's rows, with its own slice being the
    grad-carrying local tensor.  scripts/utils.py:16-23."""
    if len(local_reps) == 1:
        return local_reps[0]
    size = local_reps[0].shape[0]
    out = torch.cat([r.detach() for r in local_reps], dim=0).clone()
    out[rank * size:(rank + 1) * size] = local_reps[rank]
    return out


# ----------------------------------------------------------------------------------------------- teachers
def minmax_rows(score):
    """(s - min) / (max - min + 1e-6) per row.  bi_encoder_wrapper.py:133-137."""
    hi = score.max(dim=1, keepdim=True).values
    lo = score.min(dim=1, keepdim=True).values
    return (score - lo) / (hi - lo + 1e-6)


def ensemble_teacher_scores(teacher_q_reps, teacher_d_reps, use_in_batch_negatives=False, score_scale=30.0):
    """Average of per-teacher min-max-normalised score matrices, times score_scale.
    bi_encoder_wrapper.py:117-146 (teacher reps given; the gather is the caller's job)."""
    total = 0
    for q, d in zip(teacher_q_reps, teacher_d_reps):
        total = total + minmax_rows(student_scores(q, d, use_in_batch_negatives))
    return total / len(teacher_q_reps) * score_scale


def dense_embedding(last_hidden_state):
    """CLS vector, L2-normalised.  bi_encoder_wrapper.py:44-49."""
    cls = last_hidden_state[:, 0].float()
    return cls / cls.norm(dim=1, keepdim=True).clamp_min(1e-12)


# ----------------------------------------------------------------------------------------------- training loss
def compute_loss(q_rep, d_rep, *, loss_specs, global_step, flops_d_lambda, flops_d_T, inf_free=True,
                 flops_q_lambda=None, flops_q_T=None, flops_threshold=None, teacher_scores=None, num_processes=1):
    """Body of SparseModelTrainer.compute_loss on already gathered reps.  trainer.py:105-141.

    loss_specs: list of dicts {name, use_in_batch_negatives, temperature, weight}.
    Returns (loss * num_processes, ranking_loss, flops_loss, d_flops).
    """
    G = d_rep.shape[0] // q_rep.shape[0]
    d_flops = flops_value(d_rep, G, flops_threshold)
    flops_loss = d_flops * get_lambda(flops_d_lambda, flops_d_T, global_step)
    if not inf_free:
        flops_loss = flops_loss + flops_value(q_rep, 1, flops_threshold) * get_lambda(flops_q_lambda, flops_q_T,
                                                                                        global_step)
    rank = 0
    for spec in loss_specs:
        rank = rank + ranking_loss(spec["name"], q_rep, d_rep, teacher_scores,
                                   spec.get("use_in_batch_negatives", False), spec.get("temperature", 1.0),
                                   spec.get("weight", 1.0))
    return (rank + flops_loss) * num_processes, rank, flops_loss, d_flops


# ----------------------------------------------------------------------------------------------- encode output
def post_process(rep, id_to_token=None):
    """Dense rows -> {token: weight} dicts; column 0 never appears (it is the sentinel the reference sets to 1 and
    then drops).  sparse_encoders.py:137-150."""
    out = []
    for row in rep:
        cols = torch.nonzero(row[1:], as_tuple=True)[0] + 1
        keys = cols.tolist() if id_to_token is None else [id_to_token[c] for c in cols.tolist()]
        out.append(dict(zip(keys, row[cols].tolist())))
    return out


def document_frequency(rep):
    """count_tensor increment: number of rows with a positive weight per token.  sparse_encoders.py:178-179."""
    return (rep > 0).long().sum(dim=0)


def query_prune(token_weight_map, query_prune_ratio):
    """sparse_embedding_to_query's pruning: keep weights strictly above max * ratio.  sparse_encoders.py:184-193."""
    if query_prune_ratio <= 0:
        return dict(token_weight_map)
    limit = max(token_weight_map.values()) * query_prune_ratio
    return {k: w for k, w in token_weight_map.items() if w > limit}


def search_flops(count_q, n_q, count_d, n_d):
    """FLOPS metric of search(): dot of per-token query and doc frequencies.  scripts/search.py:82-90."""
    return float(((count_q / n_q) * (count_d / n_d)).sum())


__all__ = [name for name in dir() if not name.startswith("_") and name not in ("math", "torch")]
