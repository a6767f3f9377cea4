"""Property tests (hypothesis) of the kernels against the oracle on randomly drawn small shapes, masks and ties."""
import pytest
import torch
from hypothesis import given, settings, strategies as st

pytestmark = pytest.mark.gpu

from oracle import reference_path as R  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sparse_b200  # noqa: F401
    from sparse_b200 import ops as _ops
    return _ops


_OPS = {}


def _ops_mod():
    if "ops" not in _OPS:
        import sparse_b200  # noqa: F401
        from sparse_b200 import ops as o
        _OPS["ops"] = o
    return _OPS["ops"]


@settings(max_examples=25, deadline=None, derandomize=True)
@given(B=st.integers(1, 9), L=st.integers(1, 300), Hm=st.integers(1, 12), V=st.integers(1, 700), seed=st.integers(0, 10 ** 6),
       mask_kind=st.sampled_from(["prefix", "random", "full", "holes"]), l0=st.booleans(), quantised=st.booleans())
def test_head_forward_random_shapes_masks_and_ties(ops, B, L, Hm, V, seed, mask_kind, l0, quantised):
    H = 8 * Hm
    g = torch.Generator().manual_seed(seed)
    hidden = torch.randn(B, L, H, generator=g)
    W = torch.randn(V, H, generator=g) * 0.2
    if quantised:  # coarse values -> many exact ties between positions
        hidden = (hidden * 2).round() / 2
        W = (W * 4).round() / 4
    hidden, W = hidden.bfloat16(), W.bfloat16()
    bias = torch.randn(V, generator=g) * 0.3 - 0.2
    if mask_kind == "prefix":
        lens = torch.randint(1, L + 1, (B,), generator=g)
        mask = (torch.arange(L)[None, :] < lens[:, None]).long()
    elif mask_kind == "random":
        mask = (torch.rand(B, L, generator=g) > 0.4).long()
    elif mask_kind == "holes":
        mask = torch.ones(B, L, dtype=torch.long)
        mask[:, ::3] = 0
    else:
        mask = torch.ones(B, L, dtype=torch.long)
    o = _ops_mod()
    rep, xmax, amax = o.head_forward(hidden.cuda(), W.cuda(), bias.cuda(), mask.cuda(), use_l0=l0)
    want, values, where = R.sparse_head(hidden.float(), W.float(), bias, mask, use_l0=l0)
    torch.testing.assert_close(rep.cpu(), want, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(xmax.cpu(), values, rtol=1e-4, atol=2e-5)
    # value-only path (no arg-max requested) is bit-identical in rep
    rep2, _, _ = o.head_forward(hidden.cuda(), W.cuda(), bias.cuda(), mask.cuda(), use_l0=l0, want_aux=False)
    assert torch.equal(rep2, rep)
    # the reported position holds the maximum; with exactly representable inputs it is the first maximum
    logits = R.decoder_logits(hidden.float(), W.float(), bias) * mask.unsqueeze(-1).float()
    picked = torch.gather(logits, 1, amax.long().cpu().unsqueeze(1)).squeeze(1)
    assert bool(((values - picked).abs() <= 2e-5 * values.abs().clamp_min(1.0)).all())
    if quantised:
        active = values > 0
        assert torch.equal(amax.long().cpu()[active], where[active])


@settings(max_examples=25, deadline=None, derandomize=True)
@given(Nq=st.integers(1, 40), Lq=st.integers(1, 70), V=st.integers(8, 4000), seed=st.integers(0, 10 ** 6))
def test_idf_query_bit_exact_random(ops, Nq, Lq, V, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, V, (Nq, Lq), generator=g)
    idf = torch.rand(V, generator=g) * 6 - 1
    special = sorted(set(torch.randint(0, V, (4,), generator=g).tolist()))
    got = _ops_mod().idf_query(ids.cuda(), idf.cuda(), torch.tensor(special, dtype=torch.int32, device="cuda"))
    assert torch.equal(got.cpu(), R.idf_query(ids, idf, special))


@settings(max_examples=20, deadline=None, derandomize=True)
@given(N=st.integers(1, 12), G=st.integers(1, 5), V=st.integers(1, 1500), seed=st.integers(0, 10 ** 6),
       thr=st.one_of(st.none(), st.integers(0, 200)))
def test_flops_random(ops, N, G, V, seed, thr):
    g = torch.Generator().manual_seed(seed)
    rep = torch.relu(torch.randn(N * G, V, generator=g) - 0.5)
    x = rep.cuda().requires_grad_(True)
    val = _ops_mod().flops_value(x, G, thr)
    val.backward()
    xr = rep.clone().requires_grad_(True)
    want = R.flops_value(xr, G, thr)
    want.backward()
    torch.testing.assert_close(val.detach().cpu(), want.detach(), rtol=2e-5, atol=1e-6)
    torch.testing.assert_close(x.grad.cpu(), xr.grad, rtol=1e-4, atol=1e-7)


@settings(max_examples=20, deadline=None, derandomize=True)
@given(Nq=st.integers(1, 20), G=st.integers(1, 6), V=st.integers(2, 3000), seed=st.integers(0, 10 ** 6),
       name=st.sampled_from(["infonce", "kldiv", "marginmse"]), in_batch=st.booleans(), dense_q=st.booleans())
def test_losses_random(ops, Nq, G, V, seed, name, in_batch, dense_q):
    from sparse_b200.scripts.train.loss import LOSS_CLS_MAP
    if name != "infonce" and not in_batch and Nq == 1:
        return  # raises by contract (reference squeezes the batch dimension away)
    if name == "marginmse" and (Nq * G if in_batch else G) < 2:
        return
    g = torch.Generator().manual_seed(seed)
    keep = 0.5 if dense_q else 0.02
    q = torch.relu(torch.randn(Nq, V, generator=g)) * (torch.rand(Nq, V, generator=g) < keep)
    d = torch.relu(torch.randn(Nq * G, V, generator=g)) * (torch.rand(Nq * G, V, generator=g) < 0.3)
    teacher = torch.randn(Nq, Nq * G if in_batch else G, generator=g) * 2
    fn = LOSS_CLS_MAP[name](use_in_batch_negatives=in_batch, weight=1.3, temperature=1.7)
    qc, dc = q.cuda().requires_grad_(True), d.cuda().requires_grad_(True)
    got = fn.get_loss(qc, dc, {"scores": teacher.cuda()})
    got.backward()
    qr, dr = q.clone().requires_grad_(True), d.clone().requires_grad_(True)
    want = R.ranking_loss(name, qr, dr, teacher, in_batch, 1.7, weight=1.3)
    want.backward()
    scale = max(1.0, float(want.detach().abs()))
    torch.testing.assert_close(got.detach().cpu(), want.detach(), rtol=1e-4, atol=1e-5 * scale)
    gs = max(1e-6, float(dr.grad.abs().max()), float(qr.grad.abs().max()))
    torch.testing.assert_close(qc.grad.cpu(), qr.grad, rtol=1e-3, atol=1e-5 * gs)
    torch.testing.assert_close(dc.grad.cpu(), dr.grad, rtol=1e-3, atol=1e-5 * gs)


@settings(max_examples=15, deadline=None, derandomize=True)
@given(B=st.integers(1, 20), V=st.integers(1, 5000), seed=st.integers(0, 10 ** 6), density=st.floats(0.0, 1.0))
def test_compaction_random(ops, B, V, seed, density):
    g = torch.Generator().manual_seed(seed)
    rep = torch.rand(B, V, generator=g) * (torch.rand(B, V, generator=g) < density)
    df = torch.zeros(V, dtype=torch.int64, device="cuda")
    row_ptr, cols, vals = _ops_mod().compact_rows(rep.cuda(), first_col=1, df_count=df)
    want = R.post_process(rep)
    rp = row_ptr.cpu().tolist()
    for i, w in enumerate(want):
        assert cols[rp[i]:rp[i + 1]].cpu().tolist() == list(w.keys())
        assert vals[rp[i]:rp[i + 1]].cpu().tolist() == list(w.values())
    assert torch.equal(df.cpu(), R.document_frequency(rep))
