"""Parity of the sm_100a kernels (called through the C ABI via ops.py) against the CPU oracle and the committed
reference outputs. Tolerances are stated per test; integer/index work is bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import reference_path as R  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sparse_b200  # noqa: F401
    from sparse_b200 import ops as _ops
    return _ops


def cuda(t):
    return t.cuda() if t is not None else None


def assert_close(got, want, rtol, atol, what=""):
    torch.testing.assert_close(got.detach().float().cpu(), want.detach().float().cpu(), rtol=rtol, atol=atol, msg=lambda m: f"{what}: {m}")


def make_head_inputs(B, L, H, V, seed, ragged=True, shift=0.0, scale=0.05, mask_mode="prefix"):
    g = torch.Generator().manual_seed(seed)
    hidden = torch.randn(B, L, H, generator=g).bfloat16()
    W = (torch.randn(V, H, generator=g) * scale).bfloat16()
    bias = torch.randn(V, generator=g) * 0.1 + shift
    if mask_mode == "prefix":
        lens = torch.randint(max(1, L // 2), L + 1, (B,), generator=g) if ragged else torch.full((B,), L)
        mask = (torch.arange(L)[None, :] < lens[:, None]).long()
    elif mask_mode == "random":
        mask = (torch.rand(B, L, generator=g) > 0.3).long()
        mask[:, 0] = 1
    else:  # one fully masked row, one single-token row
        mask = torch.ones(B, L, dtype=torch.long)
        mask[0] = 0
        if B > 1:
            mask[1, 1:] = 0
    return hidden, W, bias, mask


def check_argmax(hidden, W, bias, mask, values, amax_gpu, tol=2e-5):
    """The position the kernel reports must hold the maximum (ties and rounding allowed within tol)."""
    logits = R.decoder_logits(hidden, W, bias) * mask.unsqueeze(-1).float()
    picked = torch.gather(logits, 1, amax_gpu.long().cpu().unsqueeze(1)).squeeze(1)
    scale = values.abs().clamp_min(1.0)
    assert bool(((values - picked).abs() <= tol * scale).all())


# ------------------------------------------------------------------------------------------------------ head forward
@pytest.mark.parametrize("B,L,H,V,mode", [
    (2, 128, 64, 128, "prefix"), (8, 128, 384, 30522, "prefix"), (5, 100, 384, 3000, "prefix"),
    (3, 37, 128, 1000, "random"), (4, 256, 768, 2000, "prefix"), (3, 512, 384, 1500, "prefix"),
    (2, 300, 64, 999, "random"), (3, 16, 32, 130, "edge"), (1, 1, 8, 9, "prefix"), (17, 48, 96, 700, "prefix"),
    (2, 1000, 64, 500, "prefix"),
])
@pytest.mark.parametrize("use_l0", [False, True])
def test_head_forward_vs_oracle(ops, B, L, H, V, mode, use_l0):
    hidden, W, bias, mask = make_head_inputs(B, L, H, V, seed=B * 1000 + L, shift=-0.3, mask_mode=mode)
    rep, xmax, amax = ops.head_forward(cuda(hidden), cuda(W), cuda(bias), cuda(mask), use_l0=use_l0)
    want, values, where = R.sparse_head(hidden.float(), W.float(), bias, mask, use_l0=use_l0)
    # fp32 accumulation in a different order than the CPU GEMM: rel 1e-4 / abs 2e-5 (north_star allows 1e-3)
    assert_close(rep, want, 1e-4, 2e-5, "rep")
    assert_close(xmax, values, 1e-4, 2e-5, "xmax")
    # identical arg-max up to documented tie handling: check_argmax requires EVERY reported position to hold the
    # maximum within 2e-5, so a position that differs from the oracle's can only be a value tie
    check_argmax(hidden, W, bias, mask, values, amax)


def test_head_forward_mask_dtypes_and_no_bias(ops):
    hidden, W, bias, mask = make_head_inputs(4, 64, 64, 300, seed=5)
    base, _, _ = ops.head_forward(cuda(hidden), cuda(W), cuda(bias), cuda(mask))
    for m in (mask.int(), mask.to(torch.uint8), mask.bool()):
        rep, _, _ = ops.head_forward(cuda(hidden), cuda(W), cuda(bias), cuda(m))
        assert torch.equal(rep, base)
    rep, _, _ = ops.head_forward(cuda(hidden), cuda(W), None, cuda(mask))
    want, _, _ = R.sparse_head(hidden.float(), W.float(), None, mask)
    assert_close(rep, want, 1e-4, 2e-5)


def test_head_golden_reference_outputs(ops, golden):
    for c in golden["head_hidden"]:
        for use_l0 in (False, True):
            rep = ops.sparse_head(cuda(c["hidden"]).bfloat16(), cuda(c["W"]).bfloat16(), cuda(c["bias"]), cuda(c["mask"]),
                                  use_l0=use_l0)
            assert_close(rep, c["rep"][use_l0], 1e-4, 2e-5, "golden rep")


@pytest.mark.parametrize("B,L,H,V", [(4, 128, 384, 3000), (3, 300, 768, 1200), (8, 64, 64, 500)])
def test_head_fp16_operands_vs_oracle(ops, B, L, H, V):
    """The reference configs train with `fp16: true`: the head then multiplies IEEE fp16 operands (SB200_HEAD_FP16), not
    operands re-rounded to bf16. Oracle fed the same fp16-rounded values in fp32; forward and backward."""
    g = torch.Generator().manual_seed(B + L)
    hidden = torch.randn(B, L, H, generator=g).half()
    W = (torch.randn(V, H, generator=g) * 0.1).half()
    bias = torch.randn(V, generator=g) * 0.1 - 0.3
    lens = torch.randint(L // 2, L + 1, (B,), generator=g)
    mask = (torch.arange(L)[None, :] < lens[:, None]).long()
    rep, xmax, amax = ops.head_forward(cuda(hidden), cuda(W), cuda(bias), cuda(mask), use_l0=True)
    want, values, where = R.sparse_head(hidden.float(), W.float(), bias, mask, use_l0=True)
    assert_close(rep, want, 1e-4, 2e-5, "rep (fp16 operands)")
    check_argmax(hidden, W, bias, mask, values, amax)
    # a bf16 re-rounding of the same operands would be off by ~1e-3: make sure the kernel really consumed fp16
    rep_bf16, _, _ = ops.head_forward(cuda(hidden).bfloat16(), cuda(W).bfloat16(), cuda(bias), cuda(mask), use_l0=True)
    assert float((rep_bf16.cpu() - want).abs().max()) > 10 * float((rep.cpu() - want).abs().max())
    d_rep = torch.randn(B, V, generator=g)
    dh, dw, db = ops.head_backward(cuda(d_rep), xmax, amax, cuda(hidden), cuda(W), use_l0=True)
    gh, gw, gb = R.sparse_head_grads(hidden.float(), W.float(), bias, mask, d_rep, use_l0=True)
    assert_close(dh, gh, 1e-4, 1e-5 * float(gh.abs().max()), "d_hidden (fp16 operands)")
    assert_close(dw, gw, 1e-4, 1e-5 * float(gw.abs().max()), "dW (fp16 operands)")
    assert_close(db, gb, 1e-4, 1e-5 * float(gb.abs().max()), "dbias (fp16 operands)")
    # autograd entry point: fp16 autocast keeps fp16 operands
    hc = cuda(hidden).float().requires_grad_(True)
    wc = cuda(W).float().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        out = ops.sparse_head(hc, wc, cuda(bias), cuda(mask), use_l0=True)
    assert_close(out, want, 1e-4, 2e-5, "sparse_head under fp16 autocast")


@pytest.mark.parametrize("B,L,H,V", [(5, 256, 384, 3000), (3, 512, 768, 1500), (4, 200, 64, 700), (2, 1000, 128, 600),
                                     (6, 129, 256, 900)])
def test_head_packed_input_matches_padded_and_oracle(ops, B, L, H, V):
    """Padding-free head: hidden states packed as [T, H] + cu_seqlens give the same representations, arg-max and
    gradients as the padded [B, L, H] call (and the oracle), with d_hidden returned in packed rows."""
    hidden, W, bias, mask = make_head_inputs(B, L, H, V, seed=B * 31 + L, shift=-0.3)
    mask[0] = 1                                   # one sequence without padding
    lens = mask.sum(1)
    cu = torch.zeros(B + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(lens, 0)
    filler = torch.randn(24, H).bfloat16()        # rows behind the last sequence (the packed body keeps such rows)
    packed = torch.cat([hidden[b, :lens[b]] for b in range(B)] + [filler], 0)
    assert ops.head_packed_supported(H, L)
    rep_p, xmax_p, amax_p = ops.head_forward_packed(cuda(packed), cuda(cu), L, cuda(W), cuda(bias), use_l0=True)
    rep, xmax, amax = ops.head_forward(cuda(hidden), cuda(W), cuda(bias), cuda(mask), use_l0=True)
    assert torch.equal(rep_p, rep) and torch.equal(xmax_p, xmax) and torch.equal(amax_p, amax)
    want, values, _ = R.sparse_head(hidden.float(), W.float(), bias, mask, use_l0=True)
    assert_close(rep_p, want, 1e-4, 2e-5, "packed rep vs oracle")
    g = torch.Generator().manual_seed(4)
    d_rep = torch.randn(B, V, generator=g)
    dh_p, dw_p, db_p = ops.head_backward_packed(cuda(d_rep), xmax_p, amax_p, cuda(packed), cuda(cu), L, cuda(W), use_l0=True)
    dh, dw, db = ops.head_backward(cuda(d_rep), xmax, amax, cuda(hidden), cuda(W), use_l0=True)
    gh, gw, gb = R.sparse_head_grads(hidden.float(), W.float(), bias, mask, d_rep, use_l0=True)
    assert_close(dw_p, gw, 1e-4, 1e-5 * float(gw.abs().max()), "packed dW vs oracle")
    assert_close(db_p, gb, 1e-4, 1e-5 * float(gb.abs().max()), "packed dbias vs oracle")
    assert_close(dw_p, dw, 1e-5, 1e-6 * float(gw.abs().max()), "packed dW vs padded")
    for b in range(B):
        rows = dh_p[int(cu[b]):int(cu[b + 1])]
        assert_close(rows, gh[b, :lens[b]], 1e-4, 1e-5 * float(gh.abs().max()), "packed d_hidden vs oracle")
        assert float(dh[b, lens[b]:].abs().max() if lens[b] < L else 0.0) == 0.0
    assert float(dh_p[int(cu[B]):].abs().max()) == 0.0          # filler rows receive no gradient
    # autograd entry point
    pc = cuda(packed).float().requires_grad_(True)
    wc = cuda(W).float().requires_grad_(True)
    bc = cuda(bias).requires_grad_(True)
    out = ops.sparse_head_packed(pc, cuda(cu), L, wc, bc, use_l0=True)
    (out * cuda(d_rep)).sum().backward()
    assert_close(out, want, 1e-4, 2e-5, "sparse_head_packed")
    assert_close(wc.grad, gw, 1e-4, 1e-5 * float(gw.abs().max()), "autograd dW")
    assert_close(bc.grad, gb, 1e-4, 1e-5 * float(gb.abs().max()), "autograd dbias")


def test_head_is_deterministic_and_idempotent(ops):
    hidden, W, bias, mask = make_head_inputs(6, 200, 128, 4000, seed=11)
    a = ops.head_forward(cuda(hidden), cuda(W), cuda(bias), cuda(mask))
    b = ops.head_forward(cuda(hidden), cuda(W), cuda(bias), cuda(mask))
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_head_full_size_properties(ops):
    """BASELINE sizes (160 x 256 x 384 x 30522): properties that need no CPU GEMM."""
    B, L, H, V = 160, 256, 384, 30522
    g = torch.Generator(device="cuda").manual_seed(3)
    hidden = torch.randn(B, L, H, device="cuda", generator=g).bfloat16()
    W = (torch.randn(V, H, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(V, device="cuda", generator=g) * 0.1
    lens = torch.randint(L // 2, L + 1, (B,), device="cuda", generator=g)
    mask = (torch.arange(L, device="cuda")[None, :] < lens[:, None]).long()
    rep, xmax, amax = ops.head_forward(hidden, W, bias, mask, use_l0=True)
    # (1) the reported position reproduces the reported value (checked on the GPU with an fp32 gather-dot)
    bsel = torch.arange(0, B, 7, device="cuda")
    vsel = torch.arange(0, V, 97, device="cuda")
    h = hidden[bsel][:, :, :].float()
    pos = amax[bsel][:, vsel].long()
    rows = torch.gather(h, 1, pos.unsqueeze(-1).expand(-1, -1, H))            # [b, v, H]
    dots = (rows * W[vsel].float().unsqueeze(0)).sum(-1) + bias[vsel]
    valid = torch.gather(mask[bsel], 1, pos).bool()
    dots = torch.where(valid, dots, torch.zeros_like(dots))
    torch.testing.assert_close(dots, xmax[bsel][:, vsel], rtol=1e-4, atol=2e-5)
    # (2) rep = log1p(log1p(relu(xmax))) exactly as torch computes it on the same values
    torch.testing.assert_close(rep, torch.log1p(torch.log1p(torch.relu(xmax))), rtol=2e-6, atol=1e-7)
    # (3) batch-permutation equivariance: encoding a permuted batch permutes the rows bit-exactly
    perm = torch.randperm(B, device="cuda", generator=g)
    rep2, _, _ = ops.head_forward(hidden[perm].contiguous(), W, bias, mask[perm].contiguous(), use_l0=True)
    assert torch.equal(rep2, rep[perm])
    # (4) extending the padding never changes a result (max is over real tokens, masked slots contribute 0)
    pad = 32
    hidden_p = torch.cat([hidden, torch.randn(B, pad, H, device="cuda", generator=g).bfloat16()], dim=1)
    mask_p = torch.cat([mask, torch.zeros(B, pad, dtype=torch.long, device="cuda")], dim=1)
    rep3, xmax3, _ = ops.head_forward(hidden_p, W, bias, mask_p, use_l0=True)
    full = lens == L  # rows without padding gain a masked slot -> value clamps at 0
    assert torch.equal(xmax3[~full], xmax[~full])
    assert torch.equal(xmax3[full], torch.relu(xmax[full]))


# ------------------------------------------------------------------------------------------------------ head backward
@pytest.mark.parametrize("B,L,H,V,mode,shift", [
    (3, 40, 64, 500, "prefix", 0.0), (4, 128, 384, 3000, "prefix", -0.5), (2, 300, 128, 1000, "random", 0.0),
    (5, 33, 768, 700, "prefix", -1.0), (3, 16, 32, 130, "edge", 0.0), (2, 64, 1024, 300, "prefix", 0.0),
])
@pytest.mark.parametrize("use_l0", [False, True])
def test_head_backward_vs_oracle_autograd(ops, B, L, H, V, mode, shift, use_l0):
    hidden, W, bias, mask = make_head_inputs(B, L, H, V, seed=B + L + H, shift=shift, scale=0.1, mask_mode=mode)
    g = torch.Generator().manual_seed(99)
    d_rep = torch.randn(B, V, generator=g)
    hc = cuda(hidden).requires_grad_(True)
    wc = cuda(W).float().requires_grad_(True)
    bc = cuda(bias).requires_grad_(True)
    rep = ops.sparse_head(hc, wc, bc, cuda(mask), use_l0=use_l0)
    (rep * cuda(d_rep)).sum().backward()
    gh, gw, gb = R.sparse_head_grads(hidden.float(), W.float(), bias, mask, d_rep, use_l0=use_l0)
    # d_hidden is returned in the dtype of hidden (bf16): compare with bf16 resolution; dW/dbias are fp32
    assert_close(hc.grad, gh, 1e-2, 1e-3 * float(gh.abs().max()), "d_hidden")
    assert_close(wc.grad, gw, 1e-4, 1e-5 * float(gw.abs().max() + 1), "dW")
    assert_close(bc.grad, gb, 1e-4, 1e-5 * float(gb.abs().max() + 1), "dbias")


def test_head_backward_fp32_outputs(ops):
    hidden, W, bias, mask = make_head_inputs(4, 96, 384, 2000, seed=21, shift=-0.5, scale=0.1)
    g = torch.Generator().manual_seed(7)
    d_rep = torch.randn(4, 2000, generator=g)
    rep, xmax, amax = ops.head_forward(cuda(hidden), cuda(W), cuda(bias), cuda(mask), use_l0=True)
    dh, dw, db = ops.head_backward(cuda(d_rep), xmax, amax, cuda(hidden), cuda(W), use_l0=True)
    gh, gw, gb = R.sparse_head_grads(hidden.float(), W.float(), bias, mask, d_rep, use_l0=True)
    assert_close(dh, gh, 1e-4, 1e-5 * float(gh.abs().max()), "d_hidden fp32")
    assert_close(dw, gw, 1e-4, 1e-5 * float(gw.abs().max()), "dW")
    assert_close(db, gb, 1e-4, 1e-5 * float(gb.abs().max()), "dbias")
    # linearity in d_rep
    dh2, dw2, db2 = ops.head_backward(cuda(d_rep) * 3.0, xmax, amax, cuda(hidden), cuda(W), use_l0=True)
    assert_close(dw2, dw * 3.0, 1e-5, 1e-6)
    assert_close(db2, db * 3.0, 1e-5, 1e-6)


def test_head_backward_golden_reference_grads(ops, golden):
    for c in golden["head_hidden"]:
        for use_l0 in (False, True):
            g = c["grads"][use_l0]
            rep, xmax, amax = ops.head_forward(cuda(c["hidden"]).bfloat16(), cuda(c["W"]).bfloat16(), cuda(c["bias"]),
                                               cuda(c["mask"]), use_l0=use_l0)
            dh, dw, db = ops.head_backward(cuda(g["d_rep"]), xmax, amax, cuda(c["hidden"]).bfloat16(),
                                           cuda(c["W"]).bfloat16(), use_l0=use_l0)
            assert_close(dh, g["hidden"], 1e-4, 1e-5, "golden d_hidden")
            assert_close(dw, g["W"], 1e-4, 1e-5, "golden dW")
            assert_close(db, g["bias"], 1e-4, 1e-5, "golden dbias")


def test_prune_rows(ops):
    g = torch.Generator().manual_seed(1)
    rep = torch.relu(torch.randn(7, 3001, generator=g))
    got = ops.prune_rows_(cuda(rep).clone(), 0.3)
    assert torch.equal(got.cpu(), R.prune(rep, 0.3))


# ------------------------------------------------------------------------------------------------------ inf-free query
def test_idf_query_bit_exact_golden(ops, golden):
    for c in golden["idf_query"]:
        sp = torch.tensor(c["special"], dtype=torch.int32, device="cuda")
        got = ops.idf_query(cuda(c["ids"]), cuda(c["idf"]), sp)
        assert torch.equal(got.cpu(), c["out"])


@pytest.mark.parametrize("Nq,Lq", [(32, 32), (1, 1), (7, 513), (256, 64)])
def test_idf_query_real_table(ops, idf_vector, Nq, Lq):
    g = torch.Generator().manual_seed(Nq * 31 + Lq)
    ids = torch.randint(0, 30522, (Nq, Lq), generator=g)
    ids[:, 0] = 101
    ids[:, -1] = 102
    if Lq > 4:
        ids[:, Lq // 2:] = torch.where(torch.rand(Nq, Lq - Lq // 2, generator=g) > 0.5, ids[:, Lq // 2:], torch.zeros((), dtype=torch.long))
    special = [100, 102, 0, 101, 103]
    got = ops.idf_query(cuda(ids), cuda(idf_vector), torch.tensor(special, dtype=torch.int32, device="cuda"))
    assert torch.equal(got.cpu(), R.idf_query(ids, idf_vector, special))


def test_idf_query_grad(ops):
    g = torch.Generator().manual_seed(4)
    V, Nq, Lq = 777, 9, 20
    ids = torch.randint(0, V, (Nq, Lq), generator=g)
    idf = (torch.rand(V, generator=g) * 4 - 0.3)
    w = torch.randn(Nq, V, generator=g)
    special = [0, 5]
    p = cuda(idf).requires_grad_(True)
    out = ops.idf_query(cuda(ids), p, torch.tensor(special, dtype=torch.int32, device="cuda"))
    (out * cuda(w)).sum().backward()
    pr = idf.clone().requires_grad_(True)
    present = torch.zeros(Nq, V)
    present.scatter_(1, ids, 1.0)
    present[:, special] = 0
    ((present * torch.relu(pr)) * w).sum().backward()
    assert_close(p.grad, pr.grad, 1e-6, 1e-6)


# ------------------------------------------------------------------------------------------------------ regulariser
def test_flops_golden(ops, golden):
    for c in golden["flops"]:
        got = ops.flops_value(cuda(c["rep"]), c["G"], c["thr"])
        assert_close(got, c["value"], 1e-5, 1e-6, f"flops G={c['G']} thr={c['thr']}")


@pytest.mark.parametrize("rows,G,V,thr", [(160, 5, 30522, None), (160, 5, 30522, 150), (64, 2, 30522, 9000), (30, 1, 1001, None),
                                           (512, 2, 30522, None)])
def test_flops_value_and_grad(ops, rows, G, V, thr):
    g = torch.Generator().manual_seed(rows + V)
    rep = torch.relu(torch.randn(rows, V, generator=g) - (1.0 if thr else 0.0) + torch.randn(rows, 1, generator=g) * 0.5)
    x = cuda(rep).requires_grad_(True)
    val = ops.flops_value(x, G, thr)
    (val * 0.37).backward()
    xr = rep.clone().requires_grad_(True)
    want = R.flops_value(xr, G, thr)
    (want * 0.37).backward()
    assert_close(val, want, 2e-5, 1e-6, "flops value")
    assert_close(x.grad, xr.grad, 1e-4, 1e-7, "flops grad")


# ------------------------------------------------------------------------------------------------------ scores + losses
def test_losses_golden_values_and_grads(ops, golden):
    from sparse_b200.scripts.train.loss import LOSS_CLS_MAP
    for c in golden["loss"]:
        for (name, in_batch, T), (val, gq, gd) in c["out"].items():
            fn = LOSS_CLS_MAP[name](use_in_batch_negatives=in_batch, weight=0.7, temperature=T)
            q = cuda(c["q"]).requires_grad_(True)
            d = cuda(c["d"]).requires_grad_(True)
            got = fn.get_loss(q, d, {"scores": cuda(c[("teacher", in_batch)])})
            got.backward()
            assert_close(got, val, 2e-5, 2e-6, f"{name} in_batch={in_batch} T={T}")
            assert_close(q.grad, gq, 1e-4, 1e-6, f"{name} dq")
            assert_close(d.grad, gd, 1e-4, 1e-6, f"{name} dd")


@pytest.mark.parametrize("Nq,G,V,in_batch", [(32, 5, 30522, True), (32, 5, 30522, False), (64, 2, 30522, False),
                                              (256, 2, 30522, True), (3, 1, 17, True), (70, 3, 1000, True)])
def test_scores_vs_oracle(ops, Nq, G, V, in_batch):
    g = torch.Generator().manual_seed(Nq + G)
    q = torch.relu(torch.randn(Nq, V, generator=g)) * (torch.rand(Nq, V, generator=g) > 0.9)
    d = torch.relu(torch.randn(Nq * G, V, generator=g)) * (torch.rand(Nq * G, V, generator=g) > 0.7)
    S = ops.scores(cuda(q), cuda(d), in_batch)
    want = R.student_scores(q, d, in_batch)
    assert_close(S, want, 1e-5, 1e-4, "scores")


@pytest.mark.parametrize("name", ["infonce", "kldiv", "marginmse"])
@pytest.mark.parametrize("in_batch", [False, True])
def test_loss_full_size_vs_oracle(ops, name, in_batch):
    from sparse_b200.scripts.train.loss import LOSS_CLS_MAP
    Nq, G, V = 32, 5, 30522
    g = torch.Generator().manual_seed(17)
    q = torch.relu(torch.randn(Nq, V, generator=g)) * (torch.rand(Nq, V, generator=g) > 0.999) * 3
    d = torch.relu(torch.randn(Nq * G, V, generator=g)) * (torch.rand(Nq * G, V, generator=g) > 0.99)
    teacher = torch.randn(Nq, Nq * G if in_batch else G, generator=g) * 3
    fn = LOSS_CLS_MAP[name](use_in_batch_negatives=in_batch, weight=1.0, temperature=2.0)
    qc, dc = cuda(q).requires_grad_(True), cuda(d).requires_grad_(True)
    got = fn.get_loss(qc, dc, {"scores": cuda(teacher)})
    got.backward()
    qr, dr = q.clone().requires_grad_(True), d.clone().requires_grad_(True)
    want = R.ranking_loss(name, qr, dr, teacher, in_batch, 2.0)
    want.backward()
    assert_close(got, want, 1e-4, 1e-5, name)
    assert_close(qc.grad, qr.grad, 1e-3, 1e-6, "dq")
    assert_close(dc.grad, dr.grad, 1e-3, 1e-6, "dd")


# ------------------------------------------------------------------------------------------------------ next rows
def test_compact_rows_and_df(ops, golden):
    p = golden["post"]
    rep = p["rep"]
    df = torch.zeros(rep.shape[1], dtype=torch.int64, device="cuda")
    row_ptr, cols, vals = ops.compact_rows(cuda(rep), first_col=1, df_count=df)
    row_ptr, cols, vals = row_ptr.cpu(), cols.cpu(), vals.cpu()
    want = R.post_process(rep)
    for i, w in enumerate(want):
        lo, hi = int(row_ptr[i]), int(row_ptr[i + 1])
        assert cols[lo:hi].tolist() == list(w.keys())
        assert vals[lo:hi].tolist() == list(w.values())
    assert torch.equal(df.cpu(), R.document_frequency(rep))
    g = torch.Generator().manual_seed(2)
    big = torch.relu(torch.randn(50, 30522, generator=g) - 2.0)
    row_ptr, cols, vals = ops.compact_rows(cuda(big), first_col=1)
    nz = torch.nonzero(big[:, 1:], as_tuple=True)
    assert int(row_ptr[-1]) == nz[0].numel()
    assert torch.equal(cols[: nz[0].numel()].cpu().long(), nz[1] + 1)
    assert torch.equal(vals[: nz[0].numel()].cpu(), big[:, 1:][nz])


def test_minmax_ensemble(ops, golden):
    for c in golden["ensemble"]:
        acc = None
        for q, d in zip(c["q"], c["d"]):
            S = ops.scores(cuda(q), cuda(d), c["in_batch"])
            acc = ops.minmax_accumulate(S, acc, scale=30.0 / len(c["q"]))
        assert_close(acc, c["out"], 1e-4, 1e-4, "ensemble scores")


@pytest.mark.parametrize("nnz_q", [3, 64, 511, 513, 5000])
def test_scores_sparse_and_dense_dispatch_agree(ops, nnz_q):
    """The sparse-query path (<= 512 non-zeros per query row) and the dense fallback give the same scores and
    gradients; the switch happens on the device."""
    Nq, G, V = 9, 4, 30522
    g = torch.Generator().manual_seed(nnz_q)
    q = torch.zeros(Nq, V)
    for i in range(Nq):
        cols = torch.randperm(V, generator=g)[:nnz_q if i != 2 else max(1, nnz_q // 2)]
        q[i, cols] = torch.rand(cols.numel(), generator=g) + 0.1
    d = torch.relu(torch.randn(Nq * G, V, generator=g)) * (torch.rand(Nq * G, V, generator=g) > 0.8)
    w = torch.randn(Nq, Nq * G, generator=g)
    qc, dc = cuda(q).requires_grad_(True), cuda(d).requires_grad_(True)
    S = ops.scores(qc, dc, True)
    (S * cuda(w)).sum().backward()
    qr, dr = q.clone().requires_grad_(True), d.clone().requires_grad_(True)
    Sr = qr @ dr.t()
    (Sr * w).sum().backward()
    assert_close(S, Sr, 1e-5, 1e-4, "scores")
    assert_close(dc.grad, dr.grad, 1e-5, 1e-5, "d_d")
    assert_close(qc.grad, qr.grad, 1e-5, 1e-5, "d_q")


# ------------------------------------------------------------------------------------------------------ encoder body
@pytest.mark.parametrize("R,H", [(37, 128), (4096, 384), (1000, 768), (513, 1024), (8, 256), (300, 512)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_fused_layer_norm_vs_torch(ops, R, H, dtype):
    g = torch.Generator().manual_seed(R + H)
    x = (torch.randn(R, H, generator=g) * 2 + 0.5).to(dtype)
    gamma = torch.randn(H, generator=g) * 0.5 + 1
    beta = torch.randn(H, generator=g) * 0.3
    dy = torch.randn(R, H, generator=g).to(dtype)
    xc, gc, bc = cuda(x).requires_grad_(True), cuda(gamma).requires_grad_(True), cuda(beta).requires_grad_(True)
    y = ops.layer_norm(xc, gc, bc, 1e-12)
    assert y.dtype == dtype
    y.backward(cuda(dy))
    xr, gr, br = x.float().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (H,), gr, br, 1e-12)
    yr.backward(dy.float())
    tol = dict(rtol=2e-2, atol=2e-2) if dtype == torch.bfloat16 else dict(rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(y.float().cpu(), yr, **tol)
    torch.testing.assert_close(xc.grad.float().cpu(), xr.grad, **tol)
    # parameter gradients are accumulated in fp32 from the given inputs: tight in both dtypes
    torch.testing.assert_close(gc.grad.cpu(), gr.grad, rtol=1e-4, atol=1e-3 * float(gr.grad.abs().max()))
    torch.testing.assert_close(bc.grad.cpu(), br.grad, rtol=1e-4, atol=1e-3 * float(br.grad.abs().max()))


@pytest.mark.parametrize("R,H,with_resid", [(37, 128, True), (4096, 384, True), (1000, 768, True), (513, 1024, False),
                                            (8, 256, True), (300, 512, False)])
def test_add_layer_norm_vs_torch(ops, R, H, with_resid):
    """Fused block tail without dropout against the stock autocast sequence (bf16 branch + fp32 residual -> fp32
    LayerNorm -> bf16 copy), forward and backward with gradients arriving at both copies."""
    g = torch.Generator().manual_seed(R * 7 + H)
    y = (torch.randn(R, H, generator=g) * 1.5).bfloat16()
    resid = torch.randn(R, H, generator=g) * 2 + 0.5 if with_resid else None
    gamma = torch.randn(H, generator=g) * 0.5 + 1
    beta = torch.randn(H, generator=g) * 0.3
    g32 = torch.randn(R, H, generator=g)
    g16 = torch.randn(R, H, generator=g).bfloat16()
    yc = cuda(y).requires_grad_(True)
    rc = cuda(resid).requires_grad_(True) if with_resid else None
    gc, bc = cuda(gamma).requires_grad_(True), cuda(beta).requires_grad_(True)
    o32, o16 = ops.add_layer_norm(yc, rc, gc, bc, 1e-12)
    assert o32.dtype == torch.float32 and o16.dtype == torch.bfloat16
    assert torch.equal(o16, o32.bfloat16())
    torch.autograd.backward([o32, o16], [cuda(g32), cuda(g16)])
    yr = y.float().requires_grad_(True)
    rr = resid.clone().requires_grad_(True) if with_resid else None
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    outr = torch.nn.functional.layer_norm(yr + rr if with_resid else yr, (H,), gr, br, 1e-12)
    outr.backward(g32 + g16.float())
    torch.testing.assert_close(o32.cpu(), outr, rtol=1e-5, atol=1e-5)
    assert yc.grad.dtype == torch.bfloat16
    torch.testing.assert_close(yc.grad.float().cpu(), yr.grad.bfloat16().float(), rtol=1e-2, atol=1e-5)  # 1 bf16 ulp
    if with_resid:
        torch.testing.assert_close(rc.grad.cpu(), rr.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(gc.grad.cpu(), gr.grad, rtol=1e-4, atol=1e-3 * float(gr.grad.abs().max()))
    torch.testing.assert_close(bc.grad.cpu(), br.grad, rtol=1e-4, atol=1e-3 * float(br.grad.abs().max()))
    # only one of the two copies consumed (the other gradient is absent, not zero-filled)
    for which in (0, 1):
        yc.grad = None
        outs = ops.add_layer_norm(yc, rc, gc, bc, 1e-12)
        outs[which].backward(cuda(g32) if which == 0 else cuda(g16))
        yr.grad = None
        outr = torch.nn.functional.layer_norm(yr + rr if with_resid else yr, (H,), gr, br, 1e-12)
        outr.backward(g32 if which == 0 else g16.float())
        torch.testing.assert_close(yc.grad.float().cpu(), yr.grad.bfloat16().float(), rtol=1e-2, atol=1e-5)
    # bf16-only output
    none32, only16 = ops.add_layer_norm(yc, rc, gc, bc, 1e-12, want_f32=False)
    assert none32 is None and torch.equal(only16, o16)


@pytest.mark.parametrize("p,H", [(0.1, 384), (0.5, 384), (0.1, 768), (0.3, 1024)])
def test_add_layer_norm_dropout(ops, p, H):
    """In-kernel Philox dropout: keep rate, scaling, rounding like torch's bf16 dropout, and a backward pass that
    regenerates exactly the forward mask."""
    R = 2048
    g = torch.Generator().manual_seed(5)
    y = (torch.randn(R, H, generator=g) * 0.3 + 3.0).bfloat16()    # far from zero: the mask is readable from the sum
    gamma = torch.randn(H, generator=g) * 0.5 + 1
    beta = torch.randn(H, generator=g) * 0.3
    seed = torch.tensor([123456789012345], dtype=torch.int64, device="cuda")
    yc, gc, bc = cuda(y), cuda(gamma), cuda(beta)
    o32, o16, mean, rstd = ops.add_layer_norm_forward(yc, None, gc, bc, 1e-12, seed=seed, p=p)
    s = (o32 - bc) / gc / rstd[:, None] + mean[:, None]             # dropout(y) recovered from the output
    keep = s.abs() > 0.5
    rate = float(keep.float().mean())
    assert abs(rate - (1 - p)) < 4 * (p * (1 - p) / (R * H)) ** 0.5 + 1e-3, rate
    expect = torch.where(keep, (yc.float() / (1 - p)).bfloat16().float(), torch.zeros_like(s))
    torch.testing.assert_close(s, expect, rtol=2e-3, atol=2e-3)
    # rows and columns are not correlated (a counter bug would repeat the mask)
    assert float((keep[0] == keep[1]).float().mean()) < 0.95 if p == 0.5 else True
    # same seed -> same mask; different seed -> different mask
    again = ops.add_layer_norm_forward(yc, None, gc, bc, 1e-12, seed=seed, p=p)[0]
    assert torch.equal(again, o32)
    other = ops.add_layer_norm_forward(yc, None, gc, bc, 1e-12, seed=seed + 1, p=p)[0]
    assert not torch.equal(other, o32)
    # backward against autograd through the recovered mask
    g16 = cuda(torch.randn(R, H, generator=g).bfloat16())
    yg = yc.clone().requires_grad_(True)
    out = ops.AddLayerNormFunction.apply(yg, None, gc, bc, 1e-12, seed, p, False)[1]
    out.backward(g16)
    yr = yc.float().requires_grad_(True)
    outr = torch.nn.functional.layer_norm((yr / (1 - p)).bfloat16().float() * keep, (H,), gc, bc, 1e-12)
    outr.backward(g16.float())
    assert torch.equal(yg.grad == 0, ~keep | (yr.grad == 0))
    torch.testing.assert_close(yg.grad.float(), yr.grad, rtol=2e-2, atol=2e-4)


@pytest.mark.parametrize("R,N", [(64, 8), (1000, 384), (4096, 1536), (513, 3072), (77, 4096), (300, 1152), (5, 40)])
def test_gelu_kernels_vs_torch(ops, R, N):
    """Exact-GELU forward / backward (+ fused bias gradient) against torch's erf GELU evaluated in fp32."""
    g = torch.Generator().manual_seed(R + N)
    x = (torch.randn(R, N, generator=g) * 2.5).bfloat16()
    x[0, :8] = torch.tensor([-12.0, -6.0, -4.0, -0.0, 0.0, 4.0, 6.0, 12.0]).bfloat16()
    dy = torch.randn(R, N, generator=g).bfloat16()
    xc = cuda(x)
    y = ops.gelu_forward(xc)
    xr = x.float().requires_grad_(True)
    yr = torch.nn.functional.gelu(xr)
    yr.backward(dy.float())
    # one bf16 ulp; tiny negative-tail values (|y| < 1e-4) only to absolute accuracy
    torch.testing.assert_close(y.float().cpu(), yr.detach().bfloat16().float(), rtol=8e-3, atol=2e-6)
    if R * N > 100000:   # nearly all results are bit-equal to torch's (the rest sit on a bf16 rounding boundary)
        assert float((y.float().cpu() != yr.detach().bfloat16().float()).float().mean()) < 0.03
    dx, db = ops.gelu_backward(xc, cuda(dy))
    torch.testing.assert_close(dx.float().cpu(), xr.grad.bfloat16().float(), rtol=8e-3, atol=2e-6)
    torch.testing.assert_close(db.cpu(), dx.float().sum(0).cpu(), rtol=1e-5, atol=1e-4)
    dx2, none = ops.gelu_backward(xc, cuda(dy), want_colsum=False)
    assert none is None and torch.equal(dx2, dx)


def test_linear_gelu_function(ops):
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(640, 384, generator=g)).bfloat16()
    w = (torch.randn(1536, 384, generator=g) * 0.05)
    b = torch.randn(1536, generator=g) * 0.1
    dy = torch.randn(640, 1536, generator=g).bfloat16()
    xc, wc, bc = cuda(x).requires_grad_(True), cuda(w).requires_grad_(True), cuda(b).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = ops.linear_gelu(xc, wc, bc)
    y.backward(cuda(dy))
    xr, wr, br = cuda(x).requires_grad_(True), cuda(w).requires_grad_(True), cuda(b).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        yr = torch.nn.functional.gelu(torch.nn.functional.linear(xr, wr, br))
    yr.backward(cuda(dy))
    assert y.dtype == torch.bfloat16 and wc.grad.dtype == torch.float32 and bc.grad.dtype == torch.float32
    torch.testing.assert_close(y.float(), yr.float(), rtol=1.6e-2, atol=1e-3)
    torch.testing.assert_close(xc.grad.float(), xr.grad.float(), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(wc.grad, wr.grad, rtol=2e-2, atol=5e-2)
    torch.testing.assert_close(bc.grad, br.grad, rtol=2e-2, atol=5e-2)


@pytest.mark.parametrize("n,H", [(1000, 384), (37, 64), (4096, 768)])
def test_embed_sum_vs_torch(ops, n, H):
    g = torch.Generator().manual_seed(n + H)
    nW, nP, nT = 3000, 512, 2
    W, P, T = (cuda(torch.randn(k, H, generator=g)).requires_grad_(True) for k in (nW, nP, nT))
    ids = cuda(torch.randint(0, nW, (n,), generator=g))
    ids[::5] = 7        # a hot row
    ids[1::9] = 0       # the padding row
    pos = cuda(torch.randint(0, nP, (n,), generator=g))
    typ = cuda(torch.randint(0, nT, (n,), generator=g))
    dy = cuda(torch.randn(n, H, generator=g))
    out = ops.embed_sum(ids, pos, typ, W, P, T, padding_idx=0)
    out.backward(dy)
    got = [t.grad.clone() for t in (W, P, T)]
    for t in (W, P, T):
        t.grad = None
    F = torch.nn.functional
    ref = (F.embedding(ids, W, padding_idx=0) + F.embedding(typ, T)) + F.embedding(pos, P)
    ref.backward(dy)
    assert torch.equal(out, ref)
    for a, t in zip(got, (W, P, T)):
        torch.testing.assert_close(a, t.grad, rtol=1e-4, atol=1e-4 * float(t.grad.abs().max()))
    assert float(got[0][0].abs().max()) == 0.0


def test_linear_weight_grad_is_fp32_gemm_output(ops):
    """Under bf16 autocast the weight gradient comes out of the GEMM in fp32 (no bf16 rounding of dW)."""
    g = torch.Generator().manual_seed(3)
    x = cuda(torch.randn(4096, 384, generator=g))
    w = cuda(torch.randn(384, 384, generator=g) * 0.05).requires_grad_(True)
    b = cuda(torch.randn(384, generator=g)).requires_grad_(True)
    dy = cuda(torch.randn(4096, 384, generator=g)).bfloat16()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = ops.linear(x, w, b)
    y.backward(dy)
    exact = dy.float().t() @ x.bfloat16().float()
    assert w.grad.dtype == torch.float32
    err_fp32_out = float((w.grad - exact).abs().max())
    err_bf16_out = float((exact.bfloat16().float() - exact).abs().max())
    assert err_fp32_out < 0.25 * err_bf16_out, (err_fp32_out, err_bf16_out)


def test_fused_backbone_matches_unfused(ops):
    from sparse_b200.scripts import synthetic
    V = 2000
    kw = dict(vocab_size=V, seed=1, dropout=0.0, bias_shift=-0.2)
    fused = synthetic.build_sparse_model("mini", fuse_body=True, **kw).cuda()
    plain = synthetic.build_sparse_model("mini", fuse_body=False, **kw).cuda()
    assert fused.fused_layers == 14 + 37 and plain.fused_layers == 0  # LayerNorms + Linears of the body
    assert fused.state_dict().keys() == plain.state_dict().keys()
    plain.load_state_dict(fused.state_dict())
    feats = synthetic.token_batch(6, 64, seed=2, vocab_size=V, device="cuda")
    outs = []
    for m in (fused, plain):
        m.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            rep = m(inf_free=False, **feats)
        rep.sum().backward()
        outs.append((rep.detach().float().cpu(), m.backbone.bert.embeddings.LayerNorm.weight.grad.float().cpu()))
    # bf16 activations differ in rounding points (torch normalises in fp32 and rounds later): loose tolerance
    torch.testing.assert_close(outs[0][0], outs[1][0], rtol=5e-2, atol=5e-2)
    cos = torch.nn.functional.cosine_similarity(outs[0][1], outs[1][1], dim=0)
    assert float(cos) > 0.99


@pytest.mark.parametrize("R,N", [(40960, 384), (1000, 1536), (77, 3072), (5, 8), (3000, 768), (513, 4096)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_colsum_bias_gradient(ops, R, N, dtype):
    g = torch.Generator().manual_seed(R + N)
    dy = torch.randn(R, N, generator=g).to(dtype)
    got = ops.colsum(cuda(dy))
    want = dy.double().sum(0).float()
    torch.testing.assert_close(got.cpu(), want, rtol=1e-4, atol=1e-3)


def test_fused_linear_matches_torch(ops):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(6, 50, 384, generator=g)
    w = torch.randn(1536, 384, generator=g) * 0.05
    b = torch.randn(1536, generator=g)
    dy = torch.randn(6, 50, 1536, generator=g)
    outs = []
    for fn in (ops.linear, torch.nn.functional.linear):
        xc, wc, bc = cuda(x).requires_grad_(True), cuda(w).requires_grad_(True), cuda(b).requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = fn(xc, wc, bc)
        assert y.dtype == torch.bfloat16
        y.backward(cuda(dy).bfloat16())
        outs.append((y.float().cpu(), xc.grad.cpu(), wc.grad.cpu(), bc.grad.cpu()))
    for a, r in zip(outs[0], outs[1]):
        torch.testing.assert_close(a, r, rtol=2e-2, atol=2e-2)
    # the bias gradient itself is an fp32 sum of the bf16 gradient: much tighter than torch's bf16 reduction
    torch.testing.assert_close(outs[0][3], dy.bfloat16().float().sum((0, 1)), rtol=1e-4, atol=1e-3)


def test_fused_qkv_matches_three_linears(ops):
    """One GEMM on the concatenated (cached) half weights = the three projections of transformers BertSelfAttention;
    the fp32 master parameters receive fp32 gradients (slices of one weight-gradient GEMM / one column sum)."""
    g = torch.Generator().manual_seed(4)
    x = torch.randn(300, 384, generator=g)
    ws = [torch.randn(384, 384, generator=g) * 0.05 for _ in range(3)]
    bs = [torch.randn(384, generator=g) for _ in range(3)]
    dy = torch.randn(300, 1152, generator=g)
    xc = cuda(x).requires_grad_(True)
    wc = [cuda(w).requires_grad_(True) for w in ws]
    bc = [cuda(b).requires_grad_(True) for b in bs]
    ops.refresh_half_weights(wc + bc)                      # the cached copies are what the op must pick up
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = ops.fused_qkv(xc, *wc, *bc)
    assert y.dtype == torch.bfloat16 and y.shape == (300, 1152)
    y.backward(cuda(dy).bfloat16())
    xr = cuda(x).requires_grad_(True)
    wr = [cuda(w).requires_grad_(True) for w in ws]
    br = [cuda(b).requires_grad_(True) for b in bs]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        yr = torch.cat([torch.nn.functional.linear(xr, w, b) for w, b in zip(wr, br)], 1)
    yr.backward(cuda(dy).bfloat16())
    torch.testing.assert_close(y.float(), yr.float(), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(xc.grad, xr.grad, rtol=2e-2, atol=2e-2)
    for a, r in zip(wc + bc, wr + br):
        assert a.grad.dtype == torch.float32
        torch.testing.assert_close(a.grad, r.grad, rtol=2e-2, atol=2e-2 * float(r.grad.abs().max()))
    # a parameter modified after the refresh must not be served from the stale cache
    with torch.no_grad():
        wc[0].mul_(2.0)
        y2 = ops.fused_qkv(xc.detach().bfloat16(), *wc, *bc)
    torch.testing.assert_close(y2[:, :384].float(), (2 * (y[:, :384].float() - bc[0].bfloat16().float())
                                                     + bc[0].bfloat16().float()), rtol=3e-2, atol=3e-2)


@pytest.mark.parametrize("in_batch", [False, True])
def test_local_row_backward_matches_full_backward(ops, in_batch):
    """After gather_rep only the local slice of the gathered reps carries gradient; the loss/regulariser backward
    kernels then compute just those rows. They must equal the corresponding rows of the full backward."""
    from sparse_b200.scripts.train.loss import LOSS_CLS_MAP
    Nq, G, V = 8, 3, 5000
    g = torch.Generator().manual_seed(1)
    q = torch.relu(torch.randn(Nq, V, generator=g)) * (torch.rand(Nq, V, generator=g) > 0.99)
    d = torch.relu(torch.randn(Nq * G, V, generator=g)) * (torch.rand(Nq * G, V, generator=g) > 0.9)
    teacher = torch.randn(Nq, Nq * G if in_batch else G, generator=g)
    fn = LOSS_CLS_MAP["kldiv"](use_in_batch_negatives=in_batch, temperature=2.0)
    grads = []
    for rows_q, rows_d in ((None, None), ((2, 4), (6, 12))):
        qc, dc = cuda(q).requires_grad_(True), cuda(d).requires_grad_(True)
        if rows_q is not None:
            qc._sb200_grad_rows, dc._sb200_grad_rows = rows_q, rows_d
        loss = fn.get_loss(qc, dc, {"scores": cuda(teacher)}) + 0.3 * ops.flops_value(dc, G, 20) + 0.1 * ops.flops_value(qc)
        loss.backward()
        grads.append((qc.grad.cpu(), dc.grad.cpu()))
    (fq, fd), (lq, ld) = grads
    torch.testing.assert_close(lq[2:4], fq[2:4], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(ld[6:12], fd[6:12], rtol=1e-5, atol=1e-7)
    assert float(lq[:2].abs().max()) == 0 and float(lq[4:].abs().max()) == 0
    assert float(ld[:6].abs().max()) == 0 and float(ld[12:].abs().max()) == 0
