"""The reference's UNFUSED PyTorch op sequence for the hot path, runnable on the GPU (bf16 autocast) or the CPU.

This is the same-box comparator of SURVEY.md section 2.2 / 8(d) ("the bar is the unfused PyTorch path on the same
B200") and of BASELINE.json configs[4] ("fused sparse-head microbench sweep ... vs unfused PyTorch"). It is a
measurement baseline only: nothing in the product package imports it, and it calls none of this repository's kernels.
The op sequences follow the reference line by line (paths relative to the reference root):

    sparse head          scripts/model/sparse_encoders.py:107-114 (decoder Linear inside self.backbone, `:108`)
    inf-free query       scripts/model/sparse_encoders.py:121-127
    FLOPS regulariser    scripts/train/trainer.py:61-73
    ranking losses       scripts/train/loss.py:25-43, 57-77, 86-107
    loss composition     scripts/train/trainer.py:81-143 (single process)
"""
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------- sparse head
def sparse_head(hidden, weight, bias, attention_mask, use_l0=False):
    """Decoder GEMM -> [B, L, V] logits -> mask multiply -> max over the sequence -> log1p(relu) (-> log1p)."""
    logits = F.linear(hidden, weight, bias)
    values, _ = torch.max(logits * attention_mask.unsqueeze(-1), dim=1)
    values = torch.log1p(torch.relu(values))
    return torch.log1p(values) if use_l0 else values


def encode_docs(backbone, use_l0, **features):
    """SparseModel._encode with the stock transformers backbone (its MLM head ends in the decoder Linear)."""
    logits = backbone(**features)[0]
    values, _ = torch.max(logits * features["attention_mask"].unsqueeze(-1), dim=1)
    values = torch.log1p(torch.relu(values))
    return torch.log1p(values) if use_l0 else values


def encode_inf_free(input_ids, idf_vector, special_token_ids):
    n, V = input_ids.shape[0], idf_vector.shape[0]
    out = torch.zeros(n, V, device=input_ids.device)
    out[torch.arange(n, device=input_ids.device).unsqueeze(-1), input_ids] = 1
    out[:, special_token_ids] = 0
    return out * torch.relu(idf_vector)


# ---------------------------------------------------------------------------------------------------- regulariser
def flops_value(rep, group_num=1, flops_threshold=None):
    rep = rep.reshape(-1, group_num, rep.shape[-1])
    if flops_threshold is None:
        return torch.sum(torch.mean(torch.abs(rep), dim=0) ** 2)
    w = torch.abs(rep)
    length = torch.norm(w, p=0, dim=2)
    keep = (length > flops_threshold).float().unsqueeze(2).repeat(1, 1, w.shape[2])
    return torch.sum(torch.mean(keep * w, dim=0) ** 2)


def get_lambda(value, T, step):
    return value if step >= T else value * ((step + 1) / T) ** 2


# ---------------------------------------------------------------------------------------------------- losses
def _student_scores(q_rep, d_rep, in_batch):
    if in_batch:
        return torch.matmul(q_rep, d_rep.t())
    n = q_rep.shape[0]
    docs = d_rep.reshape(n, d_rep.shape[0] // n, d_rep.shape[-1])
    return torch.bmm(docs, q_rep.reshape(n, -1, 1)).squeeze()


def infonce(q_rep, d_rep, in_batch):
    n = q_rep.shape[0]
    pos_rows = torch.arange(0, d_rep.shape[0], step=d_rep.shape[0] // n)
    scores_pos = torch.matmul(q_rep, d_rep[pos_rows].t()).diag().unsqueeze(-1)
    is_neg = torch.ones(d_rep.shape[0], dtype=torch.bool)      # built on the host every step, like the reference
    is_neg[pos_rows] = False
    neg = d_rep[is_neg]
    if in_batch:
        scores_neg = torch.matmul(q_rep, neg.t())
    else:
        scores_neg = torch.bmm(neg.reshape(n, -1, neg.shape[-1]), q_rep.reshape(n, -1, 1)).reshape(n, -1)
    scores = torch.cat([scores_pos, scores_neg], dim=1)
    target = torch.zeros(scores.shape).to(scores.device)
    target[:, 0] = 1
    return F.cross_entropy(scores, target)


def kldiv(q_rep, d_rep, teacher, in_batch, temperature=1.0):
    s = torch.log_softmax(_student_scores(q_rep, d_rep, in_batch) / temperature, dim=1)
    t = torch.softmax(teacher / temperature, dim=1)
    return F.kl_div(s, t, reduction="none").sum(dim=1).mean(dim=0)


def marginmse(q_rep, d_rep, teacher, in_batch, temperature=1.0):
    def margins(x):
        return x[:, 0].reshape(-1, 1).expand(x.shape[0], x.shape[1] - 1) - x[:, 1:]
    s = _student_scores(q_rep, d_rep, in_batch) / temperature
    return F.mse_loss(margins(s), margins(teacher / temperature))


def ranking_loss(name, q_rep, d_rep, teacher, in_batch, temperature=1.0):
    if name == "infonce":
        return infonce(q_rep, d_rep, in_batch)
    if name == "kldiv":
        return kldiv(q_rep, d_rep, teacher, in_batch, temperature)
    return marginmse(q_rep, d_rep, teacher, in_batch, temperature)


# ---------------------------------------------------------------------------------------------------- training step
class UnfusedStep:
    """One single-process training step of the reference composition (trainer.py:81-143) on stock modules:
    transformers BertForMaskedLM (untouched), the op sequences above, torch AdamW. `autocast_dtype=None` = fp32."""

    def __init__(self, backbone, idf_vector, special_token_ids, wl, autocast_dtype=torch.bfloat16, lr=2e-5,
                 weight_decay=0.01, fused_optimizer=True):
        self.backbone = backbone
        self.idf = idf_vector
        self.special = list(special_token_ids)
        self.wl = wl
        self.dtype = autocast_dtype
        dev = next(backbone.parameters()).device
        self.device = dev
        self.opt = torch.optim.AdamW(backbone.parameters(), lr=lr, weight_decay=weight_decay,
                                     fused=bool(fused_optimizer and dev.type == "cuda"))
        self.step_no = 0
        self.ema = 0.0

    def __call__(self, batch):
        wl = self.wl
        docs, queries = batch["docs"][0], batch["query"][0]
        self.backbone.train()
        ctx = torch.autocast(self.device.type, dtype=self.dtype) if self.dtype is not None else torch.autocast(
            self.device.type, enabled=False)
        with ctx:   # HF Trainer runs compute_loss inside the autocast context
            d_rep = encode_docs(self.backbone, wl["use_l0"], **docs)
            q_rep = encode_inf_free(queries["input_ids"], self.idf, self.special)
            G = d_rep.shape[0] // q_rep.shape[0]
            d_flops = flops_value(d_rep, G, wl["flops_threshold"])
            loss = d_flops * get_lambda(wl["flops_d_lambda"], wl["flops_d_T"], self.step_no)
            rank = ranking_loss(wl["loss"], q_rep, d_rep, batch.get("scores"), wl["in_batch"])
            self.ema = 0.01 * rank.item() + 0.99 * self.ema          # the reference's per-step host sync
            loss = loss + rank
        loss.backward()
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        self.step_no += 1
        return loss.detach()
