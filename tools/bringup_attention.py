"""Bring-up / A-B tool for the varlen attention kernels (csrc/attention.cu): error against an fp32 torch reference on
ragged packs (with the flash_attn library's error on the same inputs as the yardstick), dropout-mask consistency of the
forward and both backward roles, and device time against flash_attn at the C2 / C3 body shapes.
usage: python tools/bringup_attention.py [--no-time]"""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sparse_b200  # noqa: E402,F401
from sparse_b200 import _lib, ops  # noqa: E402

if "--lib" in sys.argv:      # A/B builds of the library (tools/probes/build_variants.sh); must precede the first call
    _lib.LIB_PATH = os.path.abspath(sys.argv[sys.argv.index("--lib") + 1])

try:
    from flash_attn import flash_attn_varlen_qkvpacked_func
except Exception:  # pragma: no cover
    flash_attn_varlen_qkvpacked_func = None


def reference(qkv, lens, scale, mask=None, inv_keep=1.0, dout=None):
    """fp32 per-sequence attention on the bf16 values; returns out [T, h, d] (+ dqkv when dout is given)."""
    x = qkv.float().detach().requires_grad_(dout is not None)
    outs, t0 = [], 0
    for n in lens:
        if n == 0:
            continue
        q, k, v = (x[t0:t0 + n, i].transpose(0, 1) for i in range(3))          # [h, n, d]
        p = torch.softmax(q @ k.transpose(1, 2) * scale, -1)
        if mask is not None:
            p = p * mask[:, t0:t0 + n, :n].float() * inv_keep
        outs.append((p @ v).transpose(0, 1))
        t0 += n
    out = torch.cat(outs, 0)
    if dout is None:
        return out, None
    out.backward(dout.float()[:out.shape[0]])
    return out.detach(), x.grad


def make(lens, h, d, seed=0, dev="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    T = sum(lens)
    qkv = (torch.randn(T, 3, h, d, generator=g) * 1.5).to(torch.bfloat16).to(dev)
    dout = torch.randn(T, h, d, generator=g).to(torch.bfloat16).to(dev)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    return qkv, dout, cu


def relerr(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6))


def check(lens, h, d, p_drop, tag):
    qkv, dout, cu = make(lens, h, d)
    T, L = qkv.shape[0], max(lens)
    scale = 1.0 / math.sqrt(d)
    seed = torch.tensor([0x1234567 + 77 * d], dtype=torch.int64, device=qkv.device) if p_drop > 0 else None
    out, lse = ops.attn_forward(qkv, cu, L, scale, p_drop, seed, salt=3)
    dqkv = ops.attn_backward(qkv, out, dout, lse, cu, L, scale, p_drop, seed, salt=3)
    mask, inv_keep = None, 1.0
    if p_drop > 0:
        mask = ops.attn_dropout_mask(cu, L, T, h, p_drop, seed, salt=3)
        thr = int((1 - p_drop) * 256 + 0.5)
        inv_keep = 256.0 / thr
        valid = torch.zeros(T, L, dtype=torch.bool, device=qkv.device)
        t0 = 0
        for n in lens:
            valid[t0:t0 + n, :n] = True
            t0 += n
        rate = float(mask[:, valid].float().mean())
        print(f"   keep rate {rate:.4f} (target {thr / 256:.4f})")
    want, wgrad = reference(qkv, lens, scale, mask, inv_keep, dout)
    lse_ref = []
    t0 = 0
    for n in lens:
        if n:
            q, k = qkv[t0:t0 + n, 0].float().transpose(0, 1), qkv[t0:t0 + n, 1].float().transpose(0, 1)
            lse_ref.append(torch.logsumexp(q @ k.transpose(1, 2) * scale, -1))
        t0 += n
    lse_ref = torch.cat(lse_ref, 1)
    errs = {"out": relerr(out, want), "lse": float((lse - lse_ref).abs().max()),
            "dq": relerr(dqkv[:, 0], wgrad[:, 0]), "dk": relerr(dqkv[:, 1], wgrad[:, 1]),
            "dv": relerr(dqkv[:, 2], wgrad[:, 2])}
    line = " ".join(f"{k}={v:.2e}" for k, v in errs.items())
    if flash_attn_varlen_qkvpacked_func is not None and p_drop == 0:
        x = qkv.clone().requires_grad_(True)
        fo = flash_attn_varlen_qkvpacked_func(x, cu, L, dropout_p=0.0, softmax_scale=scale, causal=False)
        fo.backward(dout)
        line += "   | flash_attn: " + " ".join(f"{k}={v:.2e}" for k, v in {
            "out": relerr(fo, want), "dq": relerr(x.grad[:, 0], wgrad[:, 0]), "dk": relerr(x.grad[:, 1], wgrad[:, 1]),
            "dv": relerr(x.grad[:, 2], wgrad[:, 2])}.items())
    bad = any(not math.isfinite(v) or v > 3e-2 for v in errs.values())
    print(f"[{'FAIL' if bad else 'ok'}] {tag}: d={d} h={h} p={p_drop} lens={lens[:8]}{'...' if len(lens) > 8 else ''}  {line}")
    if bad:
        for name, a, b in (("out", out, want), ("dq", dqkv[:, 0], wgrad[:, 0]), ("dk", dqkv[:, 1], wgrad[:, 1]),
                           ("dv", dqkv[:, 2], wgrad[:, 2])):
            diff = (a.float() - b.float()).abs().amax(dim=(1, 2))
            rows = torch.nonzero(diff > 3e-2 * b.abs().max()).flatten()[:12].tolist()
            print(f"     {name}: first bad rows {rows}; nan={int(torch.isnan(a.float()).sum())}")
    return not bad


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def bench(nseq, lo, hi, h, d, p_drop, tag):
    g = torch.Generator().manual_seed(5)
    lens = torch.randint(lo, hi + 1, (nseq,), generator=g).tolist()
    qkv, dout, cu = make(lens, h, d)
    L, scale = hi, 1.0 / math.sqrt(d)
    seed = torch.tensor([99], dtype=torch.int64, device=qkv.device) if p_drop > 0 else None
    out, lse = ops.attn_forward(qkv, cu, L, scale, p_drop, seed)
    elems = sum(n * n for n in lens) * h
    t_f = timeit(lambda: ops.attn_forward(qkv, cu, L, scale, p_drop, seed))
    t_b = timeit(lambda: ops.attn_backward(qkv, out, dout, lse, cu, L, scale, p_drop, seed))
    clk = 148 * 1.9e9
    line = (f"{tag}: own fwd {t_f:7.1f} us ({elems / t_f * 1e6 / clk:5.2f} elem/clk/SM)  bwd {t_b:7.1f} us "
            f"({elems / t_b * 1e6 / clk:5.2f})")
    if flash_attn_varlen_qkvpacked_func is not None:
        x = qkv.clone().requires_grad_(True)
        ff = lambda: flash_attn_varlen_qkvpacked_func(x, cu, L, dropout_p=p_drop, softmax_scale=scale, causal=False)  # noqa: E731
        fo = ff()
        t_ff = timeit(ff)
        t_fb = timeit(lambda: torch.autograd.grad(fo, x, dout, retain_graph=True))
        line += f"   | flash_attn fwd {t_ff:7.1f} us  bwd {t_fb:7.1f} us   -> x{t_ff / t_f:.2f} fwd, x{t_fb / t_b:.2f} bwd"
    print(line)


def mask_stats():
    """Keep rate and lag correlations of the positional dropout mask (one long sequence per head)."""
    L, h, p_drop = 512, 8, 0.1
    cu = torch.tensor([0, L, 2 * L], dtype=torch.int32, device="cuda")
    seed = torch.tensor([20260101], dtype=torch.int64, device="cuda")
    m = ops.attn_dropout_mask(cu, L, 2 * L, h, p_drop, seed, salt=1).float()      # [h, 2L, L]
    x = m - m.mean()
    var = float((x * x).mean())
    out = [f"keep {float(m.mean()):.4f}"]
    for name, a, b in (("q+1", x[:, 1:], x[:, :-1]), ("k+1", x[:, :, 1:], x[:, :, :-1]), ("q+8", x[:, 8:], x[:, :-8]),
                       ("k+8", x[:, :, 8:], x[:, :, :-8]), ("q+16", x[:, 16:], x[:, :-16]),
                       ("k+16", x[:, :, 16:], x[:, :, :-16]), ("head+1", x[1:], x[:-1]),
                       ("seq+1", x[:, L:], x[:, :L])):
        out.append(f"{name} {float((a * b).mean()) / var:+.4f}")
    rows = m.mean(2)
    cols = m[:, :L].mean(1)
    out.append(f"row-rate sd {float(rows.std()):.4f} col-rate sd {float(cols.std()):.4f} (binomial {math.sqrt(0.09 / L):.4f})")
    m2 = ops.attn_dropout_mask(cu, L, 2 * L, h, p_drop, seed + 1, salt=1).float()
    out.append(f"other seed agreement {float((m == m2).float().mean()):.4f} (independent: {0.8984 ** 2 + 0.1016 ** 2:.4f})")
    print("mask statistics: " + "  ".join(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-time", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--lib", default=None)
    args = ap.parse_args()
    torch.manual_seed(0)
    ok = True
    ragged = [256, 128, 1, 63, 64, 65, 200, 0, 37, 16, 15, 17]
    for p in (() if args.no_check else (0.0, 0.1)):
        ok &= check([64], 1, 32, p, "one tile")
        ok &= check([128, 80], 2, 32, p, "two blocks")
        ok &= check(ragged, 12, 32, p, "ragged d32")
        ok &= check([512, 300, 17, 129, 0, 64], 4, 64, p, "ragged d64")
    print("ALL OK" if ok else "FAILURES")
    mask_stats()
    if not args.no_time:
        for p in (0.0, 0.1):
            bench(160, 128, 256, 12, 32, p, f"C2 body layer (160 seqs 128..256, h12 d32, p={p})")
            bench(64, 256, 512, 12, 64, p, f"C3 body layer (64 seqs 256..512, h12 d64, p={p})")
            bench(36, 256, 512, 12, 32, p, f"C4 student layer (36 seqs 256..512, h12 d32, p={p})")


if __name__ == "__main__":
    main()
