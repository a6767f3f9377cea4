"""Pins the CPU oracle against outputs of the reference's own modules (tests/golden, made by oracle/make_golden.py)
and against the known-answer values derived in SURVEY.md section 4."""
import json
import os

import pytest
import torch

from oracle import reference_path as R

from conftest import GOLDEN


def close(a, b, rtol=1e-6, atol=1e-6):
    torch.testing.assert_close(torch.as_tensor(a, dtype=torch.float32), torch.as_tensor(b, dtype=torch.float32),
                               rtol=rtol, atol=atol)


def test_head_from_logits(golden):
    for case in golden["head"]:
        values, where = R.pooled_logits(case["logits"], case["mask"])
        for (use_l0, prune_ratio), expect in case["out"].items():
            got = R.prune(R.activation(values, use_l0), prune_ratio)
            close(got, expect, rtol=1e-6, atol=0)
        # argmax is where the reference's gradient lands
        d = case["dlogits_l0"]
        B, L, V = d.shape
        nz = d != 0
        assert int(nz.sum(dim=1).max()) <= 1
        hit = nz.float().argmax(dim=1)
        active = nz.any(dim=1)
        assert torch.equal(hit[active], where[active])


def test_head_backward_matches_reference_autograd(golden):
    for case in golden["head"]:
        logits, mask, w = case["logits"], case["mask"], case["grad_w"]
        values, where = R.pooled_logits(logits, mask)
        r1 = torch.log1p(values.clamp_min(0))
        coef = w / (1 + r1) / (1 + values.clamp_min(0)) * (values > 0)
        dl = torch.zeros_like(logits)
        dl.scatter_(1, where.unsqueeze(1), coef.unsqueeze(1))
        dl = dl * mask.unsqueeze(-1)
        close(dl, case["dlogits_l0"], rtol=1e-5, atol=1e-7)


def test_teacher_head(golden):
    for case in golden["head"]:
        values, _ = R.pooled_logits(case["logits"], case["mask"])
        out = torch.log(1 + values.clamp_min(0))
        out[:, case["teacher_special"]] = 0
        close(out, case["teacher_out"], atol=0)


def test_head_from_hidden(golden):
    for c in golden["head_hidden"]:
        for use_l0 in (False, True):
            rep, _, _ = R.sparse_head(c["hidden"], c["W"], c["bias"], c["mask"], use_l0=use_l0)
            close(rep, c["rep"][use_l0], rtol=1e-5, atol=1e-6)
            g = c["grads"][use_l0]
            gh, gw, gb = R.sparse_head_grads(c["hidden"], c["W"], c["bias"], c["mask"], g["d_rep"], use_l0=use_l0)
            close(gh, g["hidden"], rtol=1e-5, atol=1e-6)
            close(gw, g["W"], rtol=1e-5, atol=1e-6)
            close(gb, g["bias"], rtol=1e-5, atol=1e-6)
        close(R.teacher_sparse_head(c["hidden"], c["W"], c["bias"], c["mask"], c["teacher_special"]), c["teacher_out"],
              rtol=1e-5, atol=1e-6)


def test_idf_query_bit_exact(golden):
    for case in golden["idf_query"]:
        got = R.idf_query(case["ids"], case["idf"], case["special"])
        assert torch.equal(got, case["out"])


def test_flops_and_lambda(golden):
    for case in golden["flops"]:
        close(R.flops_value(case["rep"], case["G"], case["thr"]), case["value"], rtol=1e-6)
    for step, expect in golden["lambda"]:
        assert R.get_lambda(0.05, 200, step) == pytest.approx(expect, rel=1e-12)
    # SURVEY section 4 known answers
    rep = torch.tensor([[1., 0, 2, 0], [0, 0, 2, 4], [3, 0, 0, 0], [1, 1, 1, 1]])
    assert float(R.flops_value(rep, 1)) == pytest.approx(4.75)
    assert float(R.flops_value(rep, 2)) == pytest.approx(14.0)
    assert float(R.flops_value(rep, 1, 2)) == pytest.approx(0.25)
    assert float(R.flops_value(rep, 2, 2)) == pytest.approx(1.0)
    assert R.get_lambda(0.05, 200, 0) == pytest.approx(1.25e-6)
    assert R.get_lambda(0.05, 200, 99) == pytest.approx(0.0125)


def test_losses_and_grads(golden):
    for case in golden["loss"]:
        for (name, in_batch, T), (val, gq, gd) in case["out"].items():
            q = case["q"].clone().requires_grad_(True)
            d = case["d"].clone().requires_grad_(True)
            got = R.ranking_loss(name, q, d, case[("teacher", in_batch)], in_batch, T, weight=0.7)
            close(got, val, rtol=2e-6, atol=1e-6)
            got.backward()
            close(q.grad, gq, rtol=1e-5, atol=1e-6)
            close(d.grad, gd, rtol=1e-5, atol=1e-6)
    q = torch.tensor([[1., 0, 2, 0], [0, 1, 0, 1]])
    d = torch.tensor([[1., 0, 2, 0], [0, 0, 2, 4], [3, 0, 0, 0], [1, 1, 1, 1]])
    assert float(R.infonce_loss(q, d, True)) == pytest.approx(2.2752690, rel=1e-6)
    assert float(R.infonce_loss(q, d, False)) == pytest.approx(1.2200949, rel=1e-6)
    assert float(R.kldiv_loss(q, d, torch.tensor([[3., 1, 0, 2], [0, 2, 1, 3]]), True, 2.0)) == pytest.approx(0.15451002, rel=1e-5)
    assert float(R.kldiv_loss(q, d, torch.tensor([[3., 1], [0, 2]]), False, 2.0)) == pytest.approx(0.013172321, rel=1e-5)
    assert float(R.marginmse_loss(q, d, torch.tensor([[3., 1, 0, 2], [0, 2, 1, 3]]), True)) == pytest.approx(1.5)
    assert float(R.marginmse_loss(q, d, torch.tensor([[3., 1], [0, 2]]), False)) == pytest.approx(0.5)


def test_gather_rep(golden):
    g = golden["gather"]
    for rank, (out, grad) in enumerate(g["out"]):
        local = [r.clone() for r in g["reps"]]
        local[rank].requires_grad_(True)
        got = R.gather_rep(local, rank)
        assert torch.equal(got.detach(), out)
        (got * torch.arange(got.numel()).reshape(got.shape).float()).sum().backward()
        assert torch.equal(local[rank].grad, grad)


def test_teacher_ensemble(golden):
    for case in golden["ensemble"]:
        got = R.ensemble_teacher_scores(case["q"], case["d"], case["in_batch"], 30)
        close(got, case["out"], rtol=1e-5, atol=1e-5)
    c = golden["dense_embedding"]
    close(R.dense_embedding(c["hidden"]), c["out"])


def test_compute_loss(golden):
    for case in golden["compute_loss"]:
        cfg = case["cfg"]
        specs = [dict(name=n, use_in_batch_negatives=ib, temperature=T, weight=w) for n, ib, T, w in cfg["losses"]]
        loss, rank, _, _ = R.compute_loss(case["q"], case["d"], loss_specs=specs, global_step=cfg["step"],
                                          flops_d_lambda=0.05, flops_d_T=200, inf_free=cfg["inf_free"],
                                          flops_q_lambda=0.01, flops_q_T=100, flops_threshold=cfg["thr"],
                                          teacher_scores=case["teacher"])
        close(loss, case["loss"], rtol=2e-6)
        assert 0.01 * float(rank) == pytest.approx(case["moving_avg"], rel=1e-5)
    assert float(golden["compute_loss"][0]["loss"]) == pytest.approx(2.2753391, rel=1e-6)


def test_post_processing(golden):
    p = golden["post"]
    id_to_token = [f"t{i}" for i in range(p["rep"].shape[1])]
    got = R.post_process(p["rep"], id_to_token)
    assert got == p["out"]
    assert torch.equal(R.document_frequency(p["rep"]), p["df"])
    assert R.query_prune({"a": 1.0, "b": 0.3, "c": 0.05}, 0.1) == p["pruned"]["neural_sparse"]["text_sparse"]["query_tokens"]


def test_idf_vector_fixture(idf_vector):
    probe = json.load(open(os.path.join(GOLDEN, "idf_probe.json")))
    assert idf_vector.numel() == probe["n"] == 30522
    for tok, (idx, val) in probe["tokens"].items():
        assert float(idf_vector[idx]) == pytest.approx(val, rel=1e-7), tok
    assert float(idf_vector.min()) > 0 and float(idf_vector.max()) < 16


def test_port_step_matches_reference_modules_step():
    """bench.py's CPU arm: the oracle port and the reference's own modules (when the tree is present in this
    container) run the same tiny training step to the same losses, for the three workload kinds."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import pytest
    from oracle import reference_runner as RR
    if not RR.reference_available():
        pytest.skip("reference tree not present")
    import bench
    for name, extra in (("c2", {}), ("c3", {}), ("c4", {"teachers": [("dense", "tiny"), ("sparse", "tiny")]})):
        wl = dict(bench.WORKLOADS[name])
        wl.update(shape="tiny", n_queries=3, doc_len=24, query_len=8, **extra)
        losses = {}
        for prefer in (True, False):
            step, kind, _ = RR.make_cpu_step(wl, 0.0, bench.idf_vector(), prefer_reference=prefer)
            losses[kind] = [step(), step()]
        assert set(losses) == {"reference", "port"}
        for a, b in zip(losses["reference"], losses["port"]):
            assert abs(a - b) <= 1e-5 * max(1.0, abs(a)), (name, losses)
