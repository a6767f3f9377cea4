"""CPU arm of bench.py: one full training step of the reference path on the host cores.  TEST/MEASUREMENT INFRASTRUCTURE.

Two implementations of the same step (same synthetic batch, same random-init backbone, fp32, all host threads):

* kind "reference": the reference's OWN modules, imported from ``$SB200_REFERENCE_ROOT`` (default ``/root/reference``)
  when that tree exists -- ``SparseModel`` (constructed through ``__new__`` around the offline random-init backbone,
  SURVEY.md 8c), ``ModelWrapper``, the ``LOSS_CLS_MAP`` classes and ``SparseModelTrainer.compute_loss`` called unbound
  on a namespace (accelerate is absent, so the HF Trainer object cannot be built).
* kind "port": ``oracle/reference_path.py`` arithmetic around the same backbone -- used on the GPU box, where the
  reference tree does not exist.

Only ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs and ``tests/`` import this file.
"""
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("SB200_REFERENCE_ROOT", "/root/reference")
SPECIAL_IDS = [100, 102, 0, 101, 103]   # BERT-uncased special_tokens_map order


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "scripts", "train", "trainer.py"))


def _import_reference():
    def stub(name, **attrs):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        sys.modules.setdefault(name, mod)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    stub("opensearchpy", OpenSearch=object)
    stub("beir", util=types.SimpleNamespace())
    stub("beir.util")
    stub("beir.datasets")
    stub("beir.datasets.data_loader", GenericDataLoader=object)
    import scripts.model.sparse_encoders as enc
    import scripts.train.bi_encoder_wrapper as bew
    import scripts.train.loss as loss
    import scripts.train.trainer as trainer
    return enc, loss, trainer, bew


def _pool(logits, attention_mask):
    """R.pooled_logits values (the oracle's explicit arg-max loop is for small parity cases, not for timing)."""
    return torch.max(logits * attention_mask.unsqueeze(-1), dim=1).values


def _teacher_backbones(synthetic, wl):
    """Random-init stand-ins of the kd-ensemble teachers (dense gte-large shape, sparse v1 = BERT-base MLM)."""
    import transformers
    out = []
    for kind, shape in wl.get("teachers", []):
        torch.manual_seed(100 + len(out))
        cfg = transformers.BertConfig(vocab_size=synthetic.VOCAB_SIZE, max_position_embeddings=512,
                                      **synthetic.MODEL_SHAPES[shape])
        out.append((kind, transformers.BertModel(cfg, add_pooling_layer=False) if kind == "dense"
                    else transformers.BertForMaskedLM(cfg)))
    return out


def make_cpu_step(wl, bias_shift, idf_vector, n_queries=None, prefer_reference=True, seed=99):
    """-> (step() -> float loss, kind, description). One call of step() = forward + loss + backward + AdamW."""
    from sparse_b200.scripts import synthetic
    from . import reference_path as R

    torch.set_num_threads(os.cpu_count() or 1)
    nq = wl["n_queries"] if n_queries is None else n_queries
    backbone = synthetic.build_backbone(wl["shape"])
    if bias_shift:
        with torch.no_grad():
            backbone.cls.predictions.decoder.bias.add_(bias_shift)
    n_teach = len(wl.get("teachers", []))
    own_scores = wl["loss"] != "infonce" and n_teach == 0
    batch = synthetic.train_batch(nq, wl["docs_per_query"], wl["doc_len"], wl["query_len"], seed=seed,
                                  with_scores=wl["docs_per_query"] if own_scores else None, n_feature_sets=1 + n_teach)
    opt = torch.optim.AdamW(backbone.parameters(), lr=2e-5, weight_decay=0.01)
    teachers = _teacher_backbones(synthetic, wl)
    for _, t in teachers:
        t.eval()
    state = types.SimpleNamespace(global_step=0)

    if prefer_reference and reference_available():
        enc, loss_mod, trainer_mod, bew = _import_reference()
        sm = enc.SparseModel.__new__(enc.SparseModel)
        torch.nn.Module.__init__(sm)
        sm.backbone = backbone
        sm.vocab_size = synthetic.VOCAB_SIZE
        sm.special_token_ids = list(SPECIAL_IDS)
        sm.idf_vector = torch.nn.Parameter(idf_vector.clone().float(), requires_grad=False)
        sm.prune_ratio = None
        sm.use_l0 = wl["use_l0"]
        wrapper = trainer_mod.ModelWrapper(sm, True)
        fns = [loss_mod.LOSS_CLS_MAP[wl["loss"]](use_in_batch_negatives=wl["in_batch"], weight=1, temperature=1.0)]
        ns = types.SimpleNamespace(
            data_args=types.SimpleNamespace(flops_threshold=wl["flops_threshold"], flops_d_lambda=wl["flops_d_lambda"],
                                            flops_d_T=wl["flops_d_T"], flops_q_lambda=None, flops_q_T=None),
            model_args=types.SimpleNamespace(inf_free=True), loss_functions=fns, state=state,
            args=types.SimpleNamespace(logging_steps=10 ** 9),
            accelerator=types.SimpleNamespace(num_processes=1, local_process_index=0), ranking_loss_moving_avg=0)
        ns.flops_value = types.MethodType(trainer_mod.SparseModelTrainer.flops_value, ns)
        ns.get_lambda = types.MethodType(trainer_mod.SparseModelTrainer.get_lambda, ns)
        if teachers:
            wrap = bew.BiEncoderWrapper.__new__(bew.BiEncoderWrapper)
            wrap.score_scale = 30
            wrap.use_in_batch_negatives = wl["in_batch"]
            wrap.accelerator = ns.accelerator
            wrap.models = []
            for kind, bb in teachers:
                cls = bew.DenseModel if kind == "dense" else bew.BiSparseModel
                m = cls.__new__(cls)
                torch.nn.Module.__init__(m)
                m.backbone = bb
                if kind != "dense":
                    m.special_token_ids = list(SPECIAL_IDS)
                wrap.models.append(m)
            ns.bi_encoder_teacher = wrap

        def step():
            inputs = {k: (list(v) if isinstance(v, list) else v) for k, v in batch.items()}
            backbone.train()
            loss = trainer_mod.SparseModelTrainer.compute_loss(ns, wrapper, inputs)
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            state.global_step += 1
            return float(loss.detach())
        return step, "reference", f"reference modules imported from {REF_ROOT} (SparseModel, ModelWrapper, " \
                                  "LOSS_CLS_MAP, SparseModelTrainer.compute_loss unbound)"

    docs, queries = batch["docs"][0], batch["query"][0]

    def teacher_scores():
        qs, ds = [], []
        with torch.no_grad():
            for i, (kind, bb) in enumerate(teachers):
                qf, df = batch["query"][1 + i], batch["docs"][1 + i]
                if kind == "dense":
                    qs.append(R.dense_embedding(bb(**qf)[0]))
                    ds.append(R.dense_embedding(bb(**df)[0]))
                else:
                    qs.append(R.activation(_pool(bb(**qf)[0], qf["attention_mask"]), False))
                    ds.append(R.activation(_pool(bb(**df)[0], df["attention_mask"]), False))
                    for t in (qs[-1], ds[-1]):
                        t[:, SPECIAL_IDS] = 0.0
        return R.ensemble_teacher_scores(qs, ds, wl["in_batch"], 30.0)

    def step():
        backbone.train()
        scores = teacher_scores() if teachers else batch.get("scores")
        values = _pool(backbone(**docs)[0], docs["attention_mask"])
        d_rep = R.activation(values, wl["use_l0"])
        q_rep = R.idf_query(queries["input_ids"], idf_vector, SPECIAL_IDS)
        loss, _, _, _ = R.compute_loss(q_rep, d_rep, loss_specs=[dict(name=wl["loss"], use_in_batch_negatives=wl["in_batch"])],
                                       global_step=state.global_step, flops_d_lambda=wl["flops_d_lambda"],
                                       flops_d_T=wl["flops_d_T"], flops_threshold=wl["flops_threshold"],
                                       teacher_scores=scores)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        state.global_step += 1
        return float(loss.detach())
    return step, "port", "oracle/reference_path.py arithmetic around the transformers BERT body (reference tree absent)"
