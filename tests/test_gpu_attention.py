"""Variable-length attention kernels (csrc/attention.cu) against an fp32 torch restatement of
transformers BertSelfAttention (softmax(Q K^T / sqrt(d)) -> dropout -> V per sequence; the backbone call at
/root/reference/scripts/model/sparse_encoders.py:108). Floating-point kernel: bf16 operands, fp32 accumulation; the
tolerance is 1.5e-2 of the largest reference magnitude (bf16 has 8 mantissa bits; the probabilities and the score
gradients are rounded to bf16 before their second product, as in every flash-attention implementation)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1.5e-2


def _ops():
    import sparse_b200  # noqa: F401
    from sparse_b200 import ops
    return ops


def _make(lens, h, d, seed=0):
    g = torch.Generator().manual_seed(seed)
    T = sum(lens)
    qkv = (torch.randn(T, 3, h, d, generator=g) * 1.5).to(torch.bfloat16).cuda()
    dout = torch.randn(T, h, d, generator=g).to(torch.bfloat16).cuda()
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32).cuda()
    return qkv, dout, cu


def _reference(qkv, lens, scale, dout, mask=None, inv_keep=1.0):
    x = qkv.float().detach().requires_grad_(True)
    outs, lses, t0 = [], [], 0
    for n in lens:
        if n == 0:
            continue
        q, k, v = (x[t0:t0 + n, i].transpose(0, 1) for i in range(3))
        s = q @ k.transpose(1, 2) * scale
        p = torch.softmax(s, -1)
        lses.append(torch.logsumexp(s, -1).detach())
        if mask is not None:
            p = p * mask[:, t0:t0 + n, :n].float() * inv_keep
        outs.append((p @ v).transpose(0, 1))
        t0 += n
    out = torch.cat(outs, 0)
    out.backward(dout.float())
    return out.detach(), torch.cat(lses, 1), x.grad


def _close(got, want, what):
    err = float((got.float() - want.float()).abs().max())
    ref = float(want.float().abs().max())
    assert math.isfinite(err) and err <= TOL * ref, f"{what}: max err {err:.3e} vs reference magnitude {ref:.3e}"


RAGGED = [256, 128, 1, 63, 64, 65, 200, 0, 37, 16, 15, 17]


@pytest.mark.parametrize("lens,h,d", [([64], 1, 32), ([128, 80], 2, 32), (RAGGED, 12, 32), ([512, 300, 17, 129, 0, 64], 4, 64),
                                      ([257, 255, 448], 2, 64), ([1, 1, 2, 3], 3, 32)])
def test_forward_backward_match_fp32_reference(lens, h, d):
    ops = _ops()
    qkv, dout, cu = _make(lens, h, d)
    scale = 1.0 / math.sqrt(d)
    L = max(lens)
    out, lse = ops.attn_forward(qkv, cu, L, scale)
    dqkv = ops.attn_backward(qkv, out, dout, lse, cu, L, scale)
    want, want_lse, wgrad = _reference(qkv, lens, scale, dout)
    _close(out, want, "out")
    assert float((lse - want_lse).abs().max()) < 1e-4
    for i, name in enumerate(("dq", "dk", "dv")):
        _close(dqkv[:, i], wgrad[:, i], name)


@pytest.mark.parametrize("lens,h,d", [(RAGGED, 12, 32), ([512, 300, 17, 129, 0, 64], 4, 64)])
def test_dropout_mask_is_the_same_in_forward_and_both_backward_roles(lens, h, d):
    """The forward, the dQ role and the dK/dV role each regenerate the keep mask from registers in their own
    accumulator orientation; the mask hook materialises it once more. All four must agree: the reference run with the
    hook's mask reproduces out, dq, dk and dv."""
    ops = _ops()
    qkv, dout, cu = _make(lens, h, d, seed=3)
    T, L, scale, p = qkv.shape[0], max(lens), 1.0 / math.sqrt(d), 0.1
    seed = torch.tensor([987654321], dtype=torch.int64).cuda()
    out, lse = ops.attn_forward(qkv, cu, L, scale, p, seed, salt=5)
    dqkv = ops.attn_backward(qkv, out, dout, lse, cu, L, scale, p, seed, salt=5)
    mask = ops.attn_dropout_mask(cu, L, T, h, p, seed, salt=5)
    thr = int((1 - p) * 256 + 0.5)
    want, _, wgrad = _reference(qkv, lens, scale, dout, mask, 256.0 / thr)
    _close(out, want, "out")
    for i, name in enumerate(("dq", "dk", "dv")):
        _close(dqkv[:, i], wgrad[:, i], name)
    valid = torch.zeros(T, L, dtype=torch.bool).cuda()
    t0 = 0
    for n in lens:
        valid[t0:t0 + n, :n] = True
        t0 += n
    rate = float(mask[:, valid].float().mean())
    assert abs(rate - thr / 256) < 4e-3, rate
    # deterministic in (seed, salt); another salt or seed gives another mask
    out2, _ = ops.attn_forward(qkv, cu, L, scale, p, seed, salt=5)
    assert torch.equal(out, out2)
    assert not torch.equal(mask, ops.attn_dropout_mask(cu, L, T, h, p, seed, salt=6))
    assert not torch.equal(mask, ops.attn_dropout_mask(cu, L, T, h, p, seed + 1, salt=5))


_M32 = 0xFFFFFFFF


def _mix32(x):
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & _M32
    return x ^ (x >> 16)


def _mask_numpy(lens, heads, max_len, drop_p, seed, salt):
    """Integer restatement (numpy, uint64 arithmetic masked to 32 bits) of the positional dropout hash of
    csrc/attention.cu: seq_key -> patch_flags -> byte of the 2x2 patch {r, r+8} x {c, c+8}."""
    import numpy as np
    thr = min(256, max(128, int((1.0 - drop_p) * 256.0 + 0.5)))
    addc = ((128 - (256 - thr)) * 0x01010101) & _M32
    s = seed & 0xFFFFFFFFFFFFFFFF
    lo, hi = s & _M32, s >> 32
    T = sum(lens)
    out = np.zeros((heads, T, max_len), dtype=bool)
    t0 = 0
    for seq, n in enumerate(lens):
        if n == 0:
            continue
        q = np.arange(n, dtype=np.uint64)[:, None]
        k = np.arange(n, dtype=np.uint64)[None, :]
        idx = ((q >> 4) * 64 + (k >> 4)) * 64 + (q & 7) * 8 + (k & 7)
        ig = (idx * 0x9E3779B1) & _M32
        byte = ((q & 15) >> 3) * 2 + ((k & 15) >> 3)
        for head in range(heads):
            inner = _mix32((hi + 0x9E3779B9 * (seq * heads + head + 1)) & _M32)
            key = _mix32(lo ^ inner ^ ((salt * 0x85EBCA6B) & _M32))
            m = (np.uint64(key) ^ ig) * np.uint64(0xD6E8FEB9)                # 32 x 32 -> 64 bit product (no overflow)
            h = (m >> 32) ^ (m & _M32)
            flags = (((h & 0x7F7F7F7F) + addc) & _M32) | h
            out[head, t0:t0 + n, :n] = ((flags >> (8 * byte + 7)) & 1).astype(bool)
        t0 += n
    return out


def test_dropout_mask_is_bit_exact_against_the_integer_restatement():
    """Integer work: the keep mask of the kernels (through the mask hook, which the forward / backward tests tie to
    the kernels' own register-level masks) equals the numpy restatement bit for bit."""
    ops = _ops()
    lens, heads, L = [70, 0, 33, 128, 5], 3, 128
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32).cuda()
    for seed_val, salt, p in ((987654321, 5, 0.1), (-1234567890123, 0, 0.25), (2 ** 62 + 12345, 11, 0.5)):
        seed = torch.tensor([seed_val], dtype=torch.int64).cuda()
        got = ops.attn_dropout_mask(cu, L, sum(lens), heads, p, seed, salt=salt).cpu().numpy()
        want = _mask_numpy(lens, heads, L, p, seed_val, salt)
        assert (got == want).all(), (seed_val, salt, p, int((got != want).sum()))


def test_dropout_mask_statistics():
    """Positional hash: keep rate per row / column is binomial, neighbouring decisions are uncorrelated."""
    ops = _ops()
    L, h = 512, 8
    cu = torch.tensor([0, L, 2 * L], dtype=torch.int32).cuda()
    seed = torch.tensor([20260101], dtype=torch.int64).cuda()
    m = ops.attn_dropout_mask(cu, L, 2 * L, h, 0.1, seed, salt=1).float()
    x = m - m.mean()
    var = float((x * x).mean())
    for a, b in ((x[:, 1:], x[:, :-1]), (x[:, :, 1:], x[:, :, :-1]), (x[:, 8:], x[:, :-8]), (x[:, :, 8:], x[:, :, :-8]),
                 (x[:, 16:], x[:, :-16]), (x[:, :, 16:], x[:, :, :-16]), (x[1:], x[:-1]), (x[:, L:], x[:, :L])):
        assert abs(float((a * b).mean()) / var) < 0.02
    binom = math.sqrt(0.8984 * 0.1016 / L)
    assert float(m.mean(2).std()) < 1.2 * binom and float(m[:, :L].mean(1).std()) < 1.2 * binom


def test_autograd_function_and_rows_outside_the_sequences():
    ops = _ops()
    lens, h, d = [100, 60], 4, 32
    qkv, dout, cu = _make(lens + [40], h, d)            # 40 trailing rows belong to no sequence
    cu = cu[:3].contiguous()
    x = qkv.clone().requires_grad_(True)
    out = ops.varlen_attention(x, cu, 128)
    out[:160].backward(dout[:160])
    assert torch.all(x.grad[160:] == 0)
    want, _, wgrad = _reference(qkv[:160], lens, 1.0 / math.sqrt(d), dout[:160])
    _close(out[:160], want, "out")
    _close(x.grad[:160], wgrad, "dqkv")
    # eval mode: dropout off even when drop_p is given
    o2 = ops.varlen_attention(qkv, cu, 128, drop_p=0.1, training=False)
    assert torch.equal(o2[:160], out[:160].detach())


def test_filler_sequences_are_zero_filled_not_computed():
    """live_sequences: the trailing (filler) sequences of a packed batch get zero output and gradient rows; the real
    ones are unaffected."""
    ops = _ops()
    lens, h, d = [100, 60, 64, 37], 4, 32
    qkv, dout, cu = _make(lens, h, d, seed=9)
    scale = 1.0 / math.sqrt(d)
    full_out, full_lse = ops.attn_forward(qkv, cu, 128, scale)
    out, lse = ops.attn_forward(qkv, cu, 128, scale, live=2)
    assert torch.equal(out[:160], full_out[:160]) and torch.equal(lse[:, :160], full_lse[:, :160])
    assert torch.all(out[160:] == 0) and torch.all(lse[:, 160:] == 0)
    full_grad = ops.attn_backward(qkv, full_out, dout, full_lse, cu, 128, scale)
    grad = ops.attn_backward(qkv, out, dout, lse, cu, 128, scale, live=2)
    assert torch.equal(grad[:160], full_grad[:160]) and torch.all(grad[160:] == 0)
    x = qkv.clone().requires_grad_(True)
    y = ops.varlen_attention(x, cu, 128, covers_all_rows=True, live_sequences=2)
    y.backward(dout)
    assert torch.equal(x.grad, grad)


def test_argument_errors():
    import sparse_b200  # noqa: F401
    from sparse_b200 import _lib
    ops = _ops()
    qkv, dout, cu = _make([8], 2, 48)
    with pytest.raises(_lib.SparseB200Error):
        ops.attn_forward(qkv, cu, 8, 0.1)
    qkv, dout, cu = _make([8], 2, 32)
    with pytest.raises(_lib.SparseB200Error):
        ops.attn_forward(qkv, cu, 8, 0.1, 0.7, torch.zeros(1, dtype=torch.int64).cuda())
    with pytest.raises(TypeError):
        ops.attn_forward(qkv.float(), cu, 8, 0.1)


def test_packed_body_own_attention_matches_the_library_kernel():
    """Same packed body, attention='own' vs attention='flash' (eval mode: no dropout): hidden states agree to bf16
    rounding, and the training-mode step with dropout runs and gives finite gradients."""
    pytest.importorskip("flash_attn")
    import sparse_b200  # noqa: F401
    from sparse_b200.scripts import synthetic
    from sparse_b200.scripts.model.packed_body import PackedBertBody
    backbone = synthetic.build_backbone("mini", 30522, 0).cuda()
    batch = synthetic.train_batch(4, 2, 96, query_len=8, device="cuda")["docs"][0]
    bodies = {k: PackedBertBody(backbone.bert, 1.0, attention=k) for k in ("own", "flash")}
    backbone.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        res = {k: b(**batch) for k, b in bodies.items()}
    valid = res["own"][1][2]                       # rows of real tokens (the filler rows are zero-filled by the own kernels)
    hid = {k: r[0].float()[valid] for k, r in res.items()}
    assert torch.isfinite(res["own"][0].float()).all()
    err = float((hid["own"] - hid["flash"]).abs().max())
    assert err <= 0.06 * float(hid["flash"].abs().max()), err
    backbone.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        x, _ = bodies["own"](**batch)
    x.float().square().mean().backward()
    grads = [p.grad for p in backbone.bert.parameters() if p.grad is not None]
    assert grads and all(torch.isfinite(g).all() for g in grads)
