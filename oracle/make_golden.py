"""Generates tests/golden/*.pt by running the REFERENCE's own modules (imported from /root/reference) on seeded
inputs. Run in the build container only (the reference is not present on the GPU box); the outputs are committed.

    python oracle/make_golden.py

Missing third-party packages of the reference (opensearchpy, beir) are stubbed; `accelerate` is absent, so the
Trainer subclass cannot be constructed and its methods are called unbound on a SimpleNamespace, as SURVEY.md 8(c)
describes. Nothing here is product code.
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SB200_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def import_reference():
    sys.path.insert(0, REF)
    _stub("opensearchpy", OpenSearch=object)
    _stub("beir", util=types.SimpleNamespace())
    _stub("beir.util")
    _stub("beir.datasets")
    _stub("beir.datasets.data_loader", GenericDataLoader=object)
    import scripts.model.sparse_encoders as enc
    import scripts.train.loss as loss
    import scripts.utils as utils
    import scripts.train.bi_encoder_wrapper as bew
    import scripts.train.trainer as trainer
    return enc, loss, utils, bew, trainer


class FakeBackbone(torch.nn.Module):
    """Returns fixed logits, like AutoModelForMaskedLM(...)(**kw)[0]."""

    def __init__(self, logits):
        super().__init__()
        self.logits = logits

    def forward(self, **kwargs):
        return (self.logits,)


def make_sparse_model(enc, logits, V, special, idf_vector, prune_ratio, use_l0):
    m = enc.SparseModel.__new__(enc.SparseModel)
    torch.nn.Module.__init__(m)
    m.backbone = FakeBackbone(logits)
    m.vocab_size = V
    m.special_token_ids = special
    m.idf_vector = torch.nn.Parameter(idf_vector.clone(), requires_grad=False)
    m.prune_ratio = prune_ratio
    m.use_l0 = use_l0
    return m


def ragged_mask(gen, B, L):
    lens = torch.randint(max(1, L // 2), L + 1, (B,), generator=gen)
    lens[0] = L
    return (torch.arange(L)[None, :] < lens[:, None]).long()


def main():
    os.makedirs(OUT, exist_ok=True)
    enc, loss_mod, utils, bew, trainer = import_reference()
    gen = torch.Generator().manual_seed(20261017)
    golden = {}

    # ------------------------------------------------------------------ sparse head (from logits)
    head_cases = []
    for (B, L, V, shift) in [(3, 7, 50, 0.0), (4, 16, 130, -1.0), (2, 33, 257, -2.5), (1, 1, 9, 0.0)]:
        logits = torch.randn(B, L, V, generator=gen) * 2 + shift
        mask = ragged_mask(gen, B, L)
        if B > 1:
            mask[1, : max(1, L // 3)] = 1
            mask[1, max(1, L // 3):] = 0
        case = {"logits": logits, "mask": mask, "out": {}}
        for use_l0 in (False, True):
            for prune_ratio in (None, 0.1):
                m = make_sparse_model(enc, logits, V, [0, 1], torch.ones(V), prune_ratio, use_l0)
                case["out"][(use_l0, prune_ratio)] = m(inf_free=False, input_ids=None, attention_mask=mask).clone()
        # autograd of the reference head, for the backward kernels
        lg = logits.clone().requires_grad_(True)
        m = make_sparse_model(enc, lg, V, [0, 1], torch.ones(V), None, True)
        w = torch.randn(B, V, generator=gen)
        (m(inf_free=False, input_ids=None, attention_mask=mask) * w).sum().backward()
        case["grad_w"] = w
        case["dlogits_l0"] = lg.grad.clone()
        # teacher head
        t = bew.BiSparseModel.__new__(bew.BiSparseModel)
        torch.nn.Module.__init__(t)
        t.backbone = FakeBackbone(logits)
        t.special_token_ids = [0, 3]
        case["teacher_special"] = [0, 3]
        case["teacher_out"] = t(input_ids=None, attention_mask=mask).clone()
        head_cases.append(case)
    golden["head"] = head_cases

    # ------------------------------------------------------------------ head from hidden states (decoder GEMM in fp32)
    # bf16-representable inputs, so the CUDA kernels (bf16 operands, fp32 accumulate) see exactly these numbers
    hh = []
    for (B, L, H, V, shift) in [(3, 12, 16, 70, 0.0), (2, 40, 64, 300, -1.0), (5, 7, 8, 33, 0.5), (2, 130, 32, 150, -0.5)]:
        hidden = torch.randn(B, L, H, generator=gen).bfloat16().float()
        W = (torch.randn(V, H, generator=gen) * 0.3).bfloat16().float()
        bias = torch.randn(V, generator=gen) * 0.2 + shift
        mask = ragged_mask(gen, B, L)
        case = {"hidden": hidden, "W": W, "bias": bias, "mask": mask, "rep": {}, "grads": {}}
        for use_l0 in (False, True):
            lin = torch.nn.Linear(H, V)
            with torch.no_grad():
                lin.weight.copy_(W)
                lin.bias.copy_(bias)
            hid = hidden.clone().requires_grad_(True)
            m = make_sparse_model(enc, lin(hid), V, [0], torch.ones(V), None, use_l0)
            rep = m(inf_free=False, input_ids=None, attention_mask=mask)
            wgt = torch.randn(B, V, generator=gen)
            (rep * wgt).sum().backward()
            case["rep"][use_l0] = rep.detach().clone()
            case["grads"][use_l0] = {"d_rep": wgt, "hidden": hid.grad.clone(), "W": lin.weight.grad.clone(),
                                     "bias": lin.bias.grad.clone()}
        t = bew.BiSparseModel.__new__(bew.BiSparseModel)
        torch.nn.Module.__init__(t)
        with torch.no_grad():
            t.backbone = FakeBackbone(lin(hidden))
            t.special_token_ids = [0, 2]
            case["teacher_special"] = [0, 2]
            case["teacher_out"] = t(input_ids=None, attention_mask=mask).clone()
        hh.append(case)
    golden["head_hidden"] = hh

    # ------------------------------------------------------------------ inf-free query
    idf_cases = []
    for (Nq, Lq, V) in [(4, 9, 64), (6, 32, 500), (2, 5, 30522)]:
        ids = torch.randint(0, V, (Nq, Lq), generator=gen)
        ids[:, 0] = 3
        ids[0, 1] = ids[0, 2]
        idf = torch.rand(V, generator=gen) * 8 - 0.5  # some negative entries exercise relu
        special = [0, 3, 5]
        m = make_sparse_model(enc, None, V, special, idf, None, True)
        out = m(inf_free=True, input_ids=ids, attention_mask=torch.ones_like(ids))
        idf_cases.append({"ids": ids, "idf": idf, "special": special, "out": out.detach().clone()})
    golden["idf_query"] = idf_cases

    # ------------------------------------------------------------------ regulariser, lambda
    flops_cases = []
    self_ns = types.SimpleNamespace(data_args=types.SimpleNamespace(flops_threshold=None),
                                    state=types.SimpleNamespace(global_step=0))
    for (rows, G, V, thr) in [(4, 1, 4, None), (4, 2, 4, None), (4, 1, 4, 2), (4, 2, 4, 2), (12, 3, 97, None),
                              (12, 3, 97, 40), (20, 5, 301, 100), (8, 2, 64, 0)]:
        if V == 4:
            rep = torch.tensor([[1., 0, 2, 0], [0, 0, 2, 4], [3, 0, 0, 0], [1, 1, 1, 1]])
        else:
            rep = torch.relu(torch.randn(rows, V, generator=gen) + torch.randn(rows, 1, generator=gen) * 0.7)
        self_ns.data_args.flops_threshold = thr
        val = trainer.SparseModelTrainer.flops_value(self_ns, rep, G)
        flops_cases.append({"rep": rep, "G": G, "thr": thr, "value": val.clone()})
    golden["flops"] = flops_cases
    lam = []
    for step in (0, 99, 199, 200, 5000):
        self_ns.state.global_step = step
        lam.append((step, trainer.SparseModelTrainer.get_lambda(self_ns, 0.05, 200)))
    golden["lambda"] = lam

    # ------------------------------------------------------------------ ranking losses
    loss_cases = []
    for (Nq, G, V) in [(2, 2, 4), (5, 3, 40), (8, 5, 211), (1, 4, 17)]:
        if V == 4:
            q = torch.tensor([[1., 0, 2, 0], [0, 1, 0, 1]])
            d = torch.tensor([[1., 0, 2, 0], [0, 0, 2, 4], [3, 0, 0, 0], [1, 1, 1, 1]])
        else:
            q = torch.relu(torch.randn(Nq, V, generator=gen))
            d = torch.relu(torch.randn(Nq * G, V, generator=gen)) * 0.5
        case = {"q": q, "d": d, "G": G, "out": {}}
        for in_batch in (False, True):
            C = Nq * G if in_batch else G
            teacher = torch.randn(Nq, C, generator=gen) * 3
            if V == 4:
                teacher = torch.tensor([[3., 1, 0, 2], [0, 2, 1, 3]]) if in_batch else torch.tensor([[3., 1], [0, 2]])
            case[("teacher", in_batch)] = teacher
            for name in ("infonce", "kldiv", "marginmse"):
                for T in (1.0, 2.0):
                    if Nq == 1 and not in_batch and name != "infonce":
                        continue  # reference raises here (squeeze drops the batch dim), see SURVEY section 4
                    fn = loss_mod.LOSS_CLS_MAP[name](use_in_batch_negatives=in_batch, weight=0.7, temperature=T)
                    qq = q.clone().requires_grad_(True)
                    dd = d.clone().requires_grad_(True)
                    val = fn.get_loss(qq, dd, {"scores": teacher})
                    val.backward()
                    case["out"][(name, in_batch, T)] = (val.detach().clone(), qq.grad.clone(), dd.grad.clone())
        loss_cases.append(case)
    golden["loss"] = loss_cases

    # ------------------------------------------------------------------ gather_rep
    class FakeAccel:
        def __init__(self, reps, rank):
            self.reps, self.local_process_index, self.num_processes = reps, rank, len(reps)

        def gather(self, rep):
            return torch.cat([r.detach() for r in self.reps], dim=0)

    reps = [torch.randn(3, 11, generator=gen) for _ in range(3)]
    gathered = []
    for rank in range(3):
        local = reps[rank].clone().requires_grad_(True)
        acc = FakeAccel([r if i != rank else local for i, r in enumerate(reps)], rank)
        out = utils.gather_rep(local, acc)
        (out * torch.arange(out.numel()).reshape(out.shape).float()).sum().backward()
        gathered.append((out.detach().clone(), local.grad.clone()))
    golden["gather"] = {"reps": reps, "out": gathered}

    # ------------------------------------------------------------------ teacher ensemble scores
    class FakeTeacher:
        def __init__(self, q, d):
            self.q, self.d = q, d

        def __call__(self, **kw):
            return self.q if kw["which"] == "q" else self.d

    ens = []
    for in_batch in (False, True):
        Nq, G = 4, 3
        teachers = [FakeTeacher(torch.randn(Nq, 24, generator=gen), torch.randn(Nq * G, 24, generator=gen)),
                    FakeTeacher(torch.relu(torch.randn(Nq, 90, generator=gen)),
                                torch.relu(torch.randn(Nq * G, 90, generator=gen)))]
        wrap = bew.BiEncoderWrapper.__new__(bew.BiEncoderWrapper)
        wrap.score_scale = 30
        wrap.use_in_batch_negatives = in_batch
        wrap.models = teachers
        wrap.accelerator = types.SimpleNamespace(num_processes=1)
        out = wrap.get_scores_batch([{"which": "q"}] * 2, [{"which": "d"}] * 2)
        ens.append({"in_batch": in_batch, "q": [t.q for t in teachers], "d": [t.d for t in teachers], "out": out.clone()})
    golden["ensemble"] = ens
    cls = torch.randn(5, 4, 8, generator=gen)
    golden["dense_embedding"] = {"hidden": cls, "out": bew.DenseModel.get_dense_embedding((cls,)).clone()}

    # ------------------------------------------------------------------ compute_loss composition
    comp = []
    for cfg in [dict(step=1, inf_free=True, thr=None, losses=[("infonce", True, 1.0, 1.0)], kat=True),
                dict(step=57, inf_free=False, thr=None, losses=[("kldiv", False, 2.0, 1.0), ("marginmse", False, 1.0, 0.05)]),
                dict(step=3000, inf_free=True, thr=20, losses=[("kldiv", True, 1.0, 1.0)])]:
        if cfg.get("kat"):
            q = torch.tensor([[1., 0, 2, 0], [0, 1, 0, 1]])
            d = torch.tensor([[1., 0, 2, 0], [0, 0, 2, 4], [3, 0, 0, 0], [1, 1, 1, 1]])
        else:
            q = torch.relu(torch.randn(6, 150, generator=gen))
            d = torch.relu(torch.randn(12, 150, generator=gen))
        in_batch = cfg["losses"][0][1]
        teacher = torch.randn(q.shape[0], d.shape[0] if in_batch else d.shape[0] // q.shape[0], generator=gen) * 2
        fns = [loss_mod.LOSS_CLS_MAP[n](use_in_batch_negatives=ib, weight=w, temperature=T) for n, ib, T, w in cfg["losses"]]
        ns = types.SimpleNamespace(
            data_args=types.SimpleNamespace(flops_threshold=cfg["thr"], flops_d_lambda=0.05, flops_d_T=200,
                                            flops_q_lambda=0.01, flops_q_T=100),
            model_args=types.SimpleNamespace(inf_free=cfg["inf_free"]), loss_functions=fns,
            state=types.SimpleNamespace(global_step=cfg["step"]), args=types.SimpleNamespace(logging_steps=10 ** 9),
            accelerator=types.SimpleNamespace(num_processes=1, local_process_index=0), ranking_loss_moving_avg=0)
        ns.flops_value = types.MethodType(trainer.SparseModelTrainer.flops_value, ns)
        ns.get_lambda = types.MethodType(trainer.SparseModelTrainer.get_lambda, ns)
        model = lambda inputs, q=q, d=d: (d, q)
        inputs = {"query": [{"input_ids": None, "attention_mask": None}],
                  "docs": [{"input_ids": None, "attention_mask": None}], "scores": teacher}
        val = trainer.SparseModelTrainer.compute_loss(ns, model, inputs)
        comp.append({"cfg": cfg, "q": q, "d": d, "teacher": teacher, "loss": val.clone(),
                     "moving_avg": ns.ranking_loss_moving_avg})
    golden["compute_loss"] = comp

    # ------------------------------------------------------------------ post processor, DF count, query prune
    V = 40
    rep = torch.relu(torch.randn(5, V, generator=gen) - 0.8)
    rep[2] = 0
    tok = types.SimpleNamespace(vocab={f"t{i}": i for i in range(V)})
    pp = enc.SparsePostProcessor(tok)
    golden["post"] = {"rep": rep.clone(), "out": pp(rep.clone()), "df": (rep > 0).long().sum(dim=0),
                      "pruned": enc.sparse_embedding_to_query({"a": 1.0, "b": 0.3, "c": 0.05}, query_prune=0.1)}

    torch.save(golden, os.path.join(OUT, "reference_outputs.pt"))

    # ------------------------------------------------------------------ idf.json -> vector in vocab-id order
    with open(os.path.join(REF, "idf.json"), "rb") as f:
        raw = f.read()
    idf = json.loads(raw)
    vec = np.asarray(list(idf.values()), dtype=np.float32)  # keys are in BERT-uncased vocab-id order
    np.save(os.path.join(OUT, "idf_vector_f32.npy"), vec)
    keys = list(idf.keys())
    probe = {"sha256": hashlib.sha256(raw).hexdigest(), "n": len(keys),
             "tokens": {k: [keys.index(k), idf[k]] for k in ["[PAD]", "[CLS]", "[SEP]", "the", "neural", "##ing"]}}
    with open(os.path.join(OUT, "idf_probe.json"), "w") as f:
        json.dump(probe, f, indent=1)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
