"""End-to-end checks of the drop-in training path on the GPU: compute_loss against the oracle composition, the
reference's golden compute_loss values, and CUDA-graph replay against eager launches."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import reference_path as R  # noqa: E402


def _trainer(shape="tiny", loss="infonce", in_batch=True, use_l0=False, threshold=None, inf_free=True, V=2000, seed=0,
             capturable=False, base_lr=1e-3):
    import sparse_b200  # noqa: F401
    from sparse_b200.scripts import synthetic
    from sparse_b200.scripts.args import DataTrainingArguments, ModelArguments, TrainingArguments
    from sparse_b200.scripts.train.loss import LOSS_CLS_MAP
    from sparse_b200.scripts.train.trainer import SparseModelTrainer
    idf = torch.rand(V, generator=torch.Generator().manual_seed(3)) * 5
    model = synthetic.build_sparse_model(shape, idf_vector=idf, use_l0=use_l0, vocab_size=V, seed=seed, bias_shift=-0.1,
                                         dropout=0.0).cuda()
    margs = ModelArguments(inf_free=inf_free, use_l0=use_l0)
    dargs = DataTrainingArguments(loss_types=[loss], use_in_batch_negatives=in_batch, flops_d_lambda=0.05, flops_d_T=50,
                                  flops_q_lambda=0.02, flops_q_T=30, flops_threshold=threshold)
    targs = TrainingArguments(bf16=True, learning_rate=base_lr, logging_steps=10 ** 9, max_grad_norm=None, max_steps=100)
    lr = torch.tensor(base_lr, device="cuda") if capturable else base_lr
    opt = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=0.01, fused=True, capturable=capturable)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: min(1.0, (s + 1) / 5))
    fns = [LOSS_CLS_MAP[loss](use_in_batch_negatives=in_batch, weight=1.0, temperature=2.0)]
    return SparseModelTrainer(margs, dargs, fns, model=model, args=targs, optimizers=(opt, sched))


@pytest.mark.parametrize("loss,in_batch,use_l0,threshold,inf_free", [
    ("infonce", True, False, None, True), ("kldiv", False, True, 30, True), ("marginmse", True, False, None, False),
    ("kldiv", True, True, None, False),
])
def test_compute_loss_matches_oracle(loss, in_batch, use_l0, threshold, inf_free):
    from sparse_b200.scripts import synthetic
    V, nq, G = 2000, 4, 3
    tr = _trainer(loss=loss, in_batch=in_batch, use_l0=use_l0, threshold=threshold, inf_free=inf_free, V=V)
    tr.state.global_step = 7
    batch = synthetic.train_batch(nq, G, 40, query_len=10, vocab_size=V, device="cuda",
                                  with_scores=None if loss == "infonce" else (nq * G if in_batch else G))
    model = tr.model_wrapper.sparse_model

    def run(student):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return tr.model(student)
    got, outs = tr.compute_loss(run, dict(batch), return_outputs=True)

    # oracle on the same bf16-rounded decoder operands
    docs, queries = batch["docs"][0], batch["query"][0]
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        hd, dec = model.head_inputs(**docs)
        hq, _ = model.head_inputs(**queries)
    w = dec.weight.detach().bfloat16().float().cpu()
    b = dec.bias.detach().float().cpu()
    d_rep, _, _ = R.sparse_head(hd.bfloat16().float().cpu(), w, b, docs["attention_mask"].cpu(), use_l0=use_l0)
    if inf_free:
        q_rep = R.idf_query(queries["input_ids"].cpu(), model.idf_vector.detach().cpu(), model.special_token_ids)
    else:
        q_rep, _, _ = R.sparse_head(hq.bfloat16().float().cpu(), w, b, queries["attention_mask"].cpu(), use_l0=use_l0)
    want, _, _, _ = R.compute_loss(q_rep, d_rep, loss_specs=[dict(name=loss, use_in_batch_negatives=in_batch, temperature=2.0)],
                                   global_step=7, flops_d_lambda=0.05, flops_d_T=50, inf_free=inf_free, flops_q_lambda=0.02,
                                   flops_q_T=30, flops_threshold=threshold,
                                   teacher_scores=None if loss == "infonce" else batch["scores"].cpu())
    torch.testing.assert_close(outs["d_rep"].detach().cpu(), d_rep, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(outs["q_rep"].detach().cpu(), q_rep, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(got.detach().cpu(), want, rtol=2e-4, atol=1e-5)
    got.backward()
    assert all(p.grad is None or torch.isfinite(p.grad).all() for p in tr.model.parameters())


def test_compute_loss_golden_reference_values(golden):
    """Feeds the reference's own (q_rep, d_rep) pairs through the drop-in compute_loss."""
    import sparse_b200  # noqa: F401
    from sparse_b200.scripts.args import DataTrainingArguments, ModelArguments, TrainingArguments
    from sparse_b200.scripts.train.loss import LOSS_CLS_MAP
    from sparse_b200.scripts.train.trainer import SparseModelTrainer

    class Fixed(torch.nn.Module):
        def __init__(self, q, d):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1, device="cuda"))
            self.q, self.d = q, d

        def forward(self, inf_free=False, **kw):
            return self.q if kw["input_ids"] == "q" else self.d

    for case in golden["compute_loss"]:
        cfg = case["cfg"]
        fns = [LOSS_CLS_MAP[n](use_in_batch_negatives=ib, weight=w, temperature=T) for n, ib, T, w in cfg["losses"]]
        dargs = DataTrainingArguments(flops_threshold=cfg["thr"], flops_d_lambda=0.05, flops_d_T=200, flops_q_lambda=0.01, flops_q_T=100)
        tr = SparseModelTrainer(ModelArguments(inf_free=cfg["inf_free"]), dargs, fns, model=Fixed(case["q"].cuda(), case["d"].cuda()),
                                args=TrainingArguments(logging_steps=10 ** 9))
        tr.state.global_step = cfg["step"]
        inputs = {"query": [{"input_ids": "q", "attention_mask": None}], "docs": [{"input_ids": "d", "attention_mask": None}],
                  "scores": case["teacher"].cuda()}
        got = tr.compute_loss(tr.model, inputs)
        torch.testing.assert_close(got.cpu(), case["loss"], rtol=2e-5, atol=2e-6)
        assert tr.ranking_loss_moving_avg == pytest.approx(case["moving_avg"], rel=1e-4)


def test_cuda_graph_replay_matches_eager():
    from sparse_b200.scripts import synthetic
    V = 2000
    batches = [synthetic.train_batch(4, 3, 40, query_len=10, vocab_size=V, seed=50 + i, device="cuda") for i in range(8)]
    # AdamW moves every weight by ~lr per step whatever the gradient size, so the fp32-atomic ordering noise of the
    # backward kernels is amplified at a large lr; a small lr keeps the two trajectories comparable
    eager = _trainer(V=V, capturable=True, base_lr=1e-5)
    graph = _trainer(V=V, capturable=True, base_lr=1e-5)
    graph.model_wrapper.load_state_dict(copy.deepcopy(eager.model_wrapper.state_dict()))
    losses_e = [float(eager.training_step(dict(b))) for b in batches]
    # graph mode: its 3 warm-up steps run on batches[0]; mirror that on a third trainer for a like-for-like sequence
    ref = _trainer(V=V, capturable=True, base_lr=1e-5)
    ref.model_wrapper.load_state_dict(copy.deepcopy(graph.model_wrapper.state_dict()))
    for _ in range(3):
        ref.training_step(dict(batches[0]))
    graph.enable_cuda_graph(batches[0], warmup_steps=3)
    assert graph.state.global_step == ref.state.global_step == 3
    for b in batches[1:]:
        lg = float(graph.training_step(b))
        lr = float(ref.training_step(dict(b)))
        assert lg == pytest.approx(lr, rel=2e-3, abs=1e-4)
    assert graph.state.global_step == ref.state.global_step
    assert graph.ranking_loss_moving_avg == pytest.approx(ref.ranking_loss_moving_avg, rel=2e-3)
    assert losses_e[0] > 0


def test_train_ir_cli_synthetic_runs_and_checkpoints(tmp_path):
    """The reference command line (flags form) end to end on a tiny random-init model: 4 steps, checkpoint layout."""
    import sparse_b200  # noqa: F401
    from sparse_b200 import train_ir
    from sparse_b200.scripts import synthetic
    V = 2000
    out = tmp_path / "run"
    argv = ["--output_dir", str(out), "--data_type", "synthetic_posnegs", "--loss_types", "[infonce]", "--use_in_batch_negatives",
            "true", "--sample_num_one_query", "2", "--max_seq_length", "48", "--per_device_train_batch_size", "4",
            "--max_steps", "4", "--save_steps", "2", "--bf16", "true", "--learning_rate", "0.0001", "--flops_d_lambda", "0.05",
            "--flops_d_T", "10", "--max_grad_norm", "null", "--inf_free", "true", "--logging_steps", "2"]
    trainer = train_ir.main(argv, backbone=synthetic.build_backbone("tiny", V, dropout=0.0),
                            tokenizer=synthetic.SyntheticTokenizer(V))
    assert trainer.state.global_step == 4
    assert (out / "checkpoint-2" / "config.json").exists() and (out / "checkpoint-4" / "config.json").exists()
    assert (out / "config.yaml").exists() and (out / "train.log").exists()
    assert trainer.ranking_loss_moving_avg > 0
    assert trainer.last_stats is not None and trainer.last_stats["avg_doc_length"] > 0


def test_kd_ensemble_teachers_match_oracle():
    """BiEncoderWrapper with a sparse and a dense teacher (random-init) against the oracle's ensemble arithmetic."""
    import sparse_b200  # noqa: F401
    from sparse_b200.scripts import synthetic
    from sparse_b200.scripts.train.bi_encoder_wrapper import BiEncoderWrapper, BiSparseModel, DenseModel
    V, nq, G = 1500, 4, 3
    tok = synthetic.SyntheticTokenizer(V)
    sparse_t = BiSparseModel(None, backbone=synthetic.build_backbone("tiny", V, seed=3, dropout=0.0), tokenizer=tok).cuda().eval()
    class BagBackbone(torch.nn.Module):
        """Stand-in dense backbone with well-spread outputs (a random-init BERT gives nearly identical CLS vectors,
        which makes the min-max normalisation ill-conditioned): position 0 = mean token embedding."""

        def __init__(self):
            super().__init__()
            torch.manual_seed(5)
            self.emb = torch.nn.Embedding(V, 48)

        def forward(self, input_ids=None, attention_mask=None, **kw):
            e = self.emb(input_ids) * attention_mask.unsqueeze(-1)
            pooled = e.sum(1, keepdim=True) / attention_mask.sum(1).clamp_min(1).view(-1, 1, 1)
            return (pooled.expand(-1, input_ids.shape[1], -1),)

    dense_t = DenseModel(None, backbone=BagBackbone()).cuda().eval()
    batch = synthetic.train_batch(nq, G, 40, query_len=12, vocab_size=V, device="cuda")
    qf, df = batch["query"][0], batch["docs"][0]
    for in_batch in (False, True):
        wrap = BiEncoderWrapper(["dense", "sparse"], ["d", "s"], score_scale=30, use_in_batch_negatives=in_batch,
                                models=[dense_t, sparse_t])
        from sparse_b200.scripts.utils import DistEnv
        wrap.accelerator = DistEnv()
        got = wrap.get_scores_batch([qf, qf], [df, df])
        with torch.no_grad():
            reps_q = [dense_t(**qf).float().cpu(), sparse_t(**qf).float().cpu()]
            reps_d = [dense_t(**df).float().cpu(), sparse_t(**df).float().cpu()]
        want = R.ensemble_teacher_scores(reps_q, reps_d, in_batch, 30.0)
        torch.testing.assert_close(got.cpu(), want, rtol=1e-4, atol=1e-3)
    # the sparse teacher head itself against the oracle (single log, specials zeroed)
    with torch.no_grad():
        hidden = sparse_t._split.transform(sparse_t._split.body(**df)[0])
        dec = sparse_t._split.decoder
        want = R.teacher_sparse_head(hidden.bfloat16().float().cpu(), dec.weight.bfloat16().float().cpu(), dec.bias.float().cpu(),
                                     df["attention_mask"].cpu(), sparse_t.special_token_ids)
        torch.testing.assert_close(sparse_t(**df).cpu(), want, rtol=1e-4, atol=2e-5)


def test_sparse_encoder_encode_output_and_flops_metric():
    import sparse_b200  # noqa: F401
    from sparse_b200.scripts import synthetic
    from sparse_b200.scripts.model.sparse_encoders import SparseEncoder, sparse_embedding_to_query
    from sparse_b200.scripts.search import flops_metric
    V = 1500
    model = synthetic.build_sparse_model("tiny", vocab_size=V, bias_shift=-0.5, dropout=0.0).cuda().eval()
    enc = SparseEncoder(model, max_length=32)
    feats = synthetic.token_batch(6, 32, seed=3, vocab_size=V, device="cuda")
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        rep = model(inf_free=False, **feats)
    out = enc.encode_output(rep)
    id_to_token = enc.post_processor.id_to_token
    want = R.post_process(rep.cpu(), id_to_token)
    assert out == want
    assert torch.equal(enc.count_tensor.cpu().long(), R.document_frequency(rep.cpu()))
    q = sparse_embedding_to_query(out[0], query_prune=0.2)["neural_sparse"]["text_sparse"]["query_tokens"]
    assert q == R.query_prune(out[0], 0.2)
    fl, ql, dl = flops_metric(enc.count_tensor, 6, enc.count_tensor, 6)
    assert fl == pytest.approx(R.search_flops(enc.count_tensor.cpu(), 6, enc.count_tensor.cpu(), 6), rel=1e-6)
    assert dl == pytest.approx(float((rep > 0).sum()) / 6, rel=1e-6) and ql == dl


def test_fp16_autocast_training_step_runs():
    """The reference configs train with fp16 + GradScaler: the same path (fp16 autocast, scaled loss through the custom
    backward kernels, unscale, step) must run and produce finite parameters."""
    import sparse_b200  # noqa: F401
    from sparse_b200.scripts import synthetic
    from sparse_b200.scripts.args import DataTrainingArguments, ModelArguments, TrainingArguments
    from sparse_b200.scripts.train.loss import LOSS_CLS_MAP
    from sparse_b200.scripts.train.trainer import SparseModelTrainer
    V = 2000
    model = synthetic.build_sparse_model("mini", vocab_size=V, bias_shift=-0.1, dropout=0.1).cuda()
    targs = TrainingArguments(fp16=True, learning_rate=1e-4, logging_steps=10 ** 9, max_grad_norm=1.0, max_steps=10)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    tr = SparseModelTrainer(ModelArguments(inf_free=True), DataTrainingArguments(loss_types=["infonce"], use_in_batch_negatives=True,
                                                                                 flops_d_lambda=0.05, flops_d_T=10),
                            [LOSS_CLS_MAP["infonce"](use_in_batch_negatives=True)], model=model, args=targs, optimizers=(opt, None))
    assert tr.scaler is not None
    losses = []
    for i in range(3):
        batch = synthetic.train_batch(4, 3, 64, query_len=12, vocab_size=V, seed=200 + i, device="cuda")
        losses.append(float(tr.training_step(batch)))
    assert all(l == l and abs(l) < 1e6 for l in losses), losses
    assert all(torch.isfinite(p).all() for p in model.parameters())


@pytest.mark.parametrize("capacity,mask_kind,L", [(1.0, "prefix", 96), (0.9, "prefix", 96), (1.0, "holes", 96),
                                                    (0.9, "prefix", 192), (1.0, "holes", 192)])   # L > 128: packed head too
def test_padding_free_body_matches_padded_body(capacity, mask_kind, L):
    """The packed (unpadded, flash-attn varlen) body gives the same sparse vectors and gradients as the padded
    transformers body on the real tokens."""
    import sparse_b200  # noqa: F401
    from sparse_b200.scripts import synthetic
    V = 2000
    kw = dict(vocab_size=V, seed=1, dropout=0.0, bias_shift=-0.2)
    packed = synthetic.build_sparse_model("mini", unpad_capacity=capacity, **kw).cuda()
    padded = synthetic.build_sparse_model("mini", **kw).cuda()
    if packed.__dict__["_packed"] is None:
        pytest.skip("flash_attn varlen kernels unavailable")
    assert packed.state_dict().keys() == padded.state_dict().keys()
    padded.load_state_dict(packed.state_dict())
    feats = synthetic.token_batch(12, L, seed=2, vocab_size=V, device="cuda")
    if mask_kind == "holes":
        feats["attention_mask"][:, 5::7] = 0
    g = torch.Generator(device="cuda").manual_seed(0)
    w = torch.randn(12, V, device="cuda", generator=g)
    outs = []
    for m in (packed, padded):
        m.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            rep = m(inf_free=False, **feats)
        (rep * w).sum().backward()
        outs.append((rep.detach().float().cpu(), m.backbone.bert.encoder.layer[0].intermediate.dense.weight.grad.float().cpu(),
                     m.backbone.bert.embeddings.word_embeddings.weight.grad.float().cpu()))
    assert packed.unpad_overflows() == 0
    torch.testing.assert_close(outs[0][0], outs[1][0], rtol=5e-2, atol=5e-2)   # bf16 bodies, different attention kernels
    for a, b in zip(outs[0][1:], outs[1][1:]):
        cos = torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0)
        assert float(cos) > 0.99, float(cos)
    # a capacity that is too small is detected, not silently accepted
    tight = synthetic.build_sparse_model("mini", unpad_capacity=0.3, **kw).cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        tight(inf_free=False, **feats)
    assert tight.unpad_overflows() == 1


def test_overflowed_packed_batch_never_reaches_the_weights():
    """unpad_capacity too small for the batch: the step runs (no host sync inside it), but the fused optimizer skips the
    update on the device flag -- weights, moments and step counts untouched -- and the trainer raises afterwards."""
    import sparse_b200  # noqa: F401
    from sparse_b200.scripts import synthetic
    from sparse_b200.scripts.args import DataTrainingArguments, ModelArguments, TrainingArguments
    from sparse_b200.scripts.train.loss import LOSS_CLS_MAP
    from sparse_b200.scripts.train.trainer import SparseModelTrainer
    V = 2000
    model = synthetic.build_sparse_model("mini", vocab_size=V, bias_shift=-0.1, dropout=0.0, unpad_capacity=0.9).cuda()
    if model.__dict__["_packed"] is None:
        pytest.skip("flash_attn varlen kernels unavailable")
    targs = TrainingArguments(bf16=True, learning_rate=1e-3, logging_steps=10 ** 9, max_grad_norm=None, max_steps=10)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
    tr = SparseModelTrainer(ModelArguments(inf_free=True), DataTrainingArguments(loss_types=["infonce"], use_in_batch_negatives=True,
                                                                                 flops_d_lambda=0.05, flops_d_T=10),
                            [LOSS_CLS_MAP["infonce"](use_in_batch_negatives=True)], model=model, args=targs, optimizers=(opt, None))
    ok = synthetic.train_batch(4, 3, 192, query_len=12, vocab_size=V, seed=300, device="cuda")      # lengths in [L/2, L]
    full = synthetic.train_batch(4, 3, 192, query_len=12, vocab_size=V, seed=301, device="cuda")
    full["docs"][0]["attention_mask"][:] = 1                                                         # 100 % real tokens > 0.9
    before = [p.detach().clone() for p in model.parameters()]
    tr.training_step(ok)
    tr.check_unpad()
    after_ok = [p.detach().clone() for p in model.parameters()]
    assert any(not torch.equal(a, b) for a, b in zip(before, after_ok)), "a fitting batch must train"
    tr.training_step(full)
    torch.cuda.synchronize()
    assert all(torch.equal(a, p.detach()) for a, p in zip(after_ok, model.parameters())), "overflowed batch leaked into the weights"
    with pytest.raises(RuntimeError, match="unpad_capacity"):
        tr.check_unpad()
    with pytest.raises(RuntimeError, match="unpad_capacity"):
        tr.training_step(ok)          # the asynchronous per-step poll has seen the counter by now


def test_train_ir_on_text_rows_and_encode_texts(tmp_path):
    """The reference's `data_type: posnegs` pipeline end to end (dataset -> collator -> prefetch -> training steps) with
    the offline stand-in tokenizer, then SparseEncoder.encode on raw texts (tokenizer half included)."""
    import json
    import sparse_b200  # noqa: F401
    from sparse_b200 import train_ir
    from sparse_b200.scripts import synthetic
    from sparse_b200.scripts.model.sparse_encoders import SparseEncoder
    V = 1500
    path = tmp_path / "train.jsonl"
    with open(path, "w") as f:
        for i in range(12):
            f.write(json.dumps({"query": f"what is item {i}", "pos": f"item {i} is a thing with number {i}",
                                "negs": [f"unrelated text {i} {k} about something else entirely" for k in range(4)]}) + "\n")
    argv = ["--output_dir", str(tmp_path / "out"), "--data_type", "posnegs", "--train_file", str(path),
            "--loss_types", "[infonce]", "--use_in_batch_negatives", "true", "--sample_num_one_query", "2",
            "--max_seq_length", "24", "--per_device_train_batch_size", "4", "--max_steps", "4", "--bf16", "true",
            "--logging_steps", "2", "--save_strategy", "no", "--flops_d_lambda", "0.01", "--flops_d_T", "10",
            "--learning_rate", "1e-4", "--dataloader_drop_last", "true"]
    trainer = train_ir.main(argv, backbone=synthetic.build_backbone("tiny", V, dropout=0.0),
                            tokenizer=synthetic.SyntheticTokenizer(V))
    assert trainer.state.global_step == 4
    assert float(trainer.ranking_loss_moving_avg) > 0
    model = trainer.model_wrapper.sparse_model.eval()
    enc = SparseEncoder(model, max_length=24)
    out = enc.encode(["item 3 is a thing", "unrelated text about something"])
    assert len(out) == 2 and all(isinstance(d, dict) for d in out)
    assert all(w > 0 for d in out for w in d.values())
    q = enc.encode(["what is item 3"], inf_free=True)[0]
    assert set(q) <= {"tok%d" % i for i in range(V)} and 1 <= len(q) <= 4      # 4 words (hash collisions may merge two)
    assert all(w == 1.0 for w in q.values())                                     # default idf weight, specials dropped
