"""Launches every kernel of libsparse_b200.so once or twice on the C2 / C3 shapes, for `ncu --set full -k regex:sb200`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_b200
from sparse_b200 import ops

def run(B, L, H, V, G, nq, regime_shift):
    g = torch.Generator(device="cuda").manual_seed(0)
    hidden = torch.randn(B, L, H, device="cuda", generator=g).bfloat16()
    W = (torch.randn(V, H, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(V, device="cuda", generator=g) * 0.1 + regime_shift
    lens = torch.randint(L // 2, L + 1, (B,), device="cuda", generator=g)
    mask = (torch.arange(L, device="cuda")[None, :] < lens[:, None]).long()
    for _ in range(2):
        rep, xmax, amax = ops.head_forward(hidden, W, bias, mask, use_l0=False)
    d_rep = torch.randn(B, V, device="cuda", generator=g)
    ops.head_backward(d_rep, xmax, amax, hidden, W)
    ids = torch.randint(1000, V, (nq, 32), device="cuda", generator=g)
    idf = torch.rand(V, device="cuda", generator=g) * 8
    q = ops.idf_query_forward(ids, idf, torch.tensor([0, 100, 101, 102, 103], dtype=torch.int32, device="cuda"))
    ops.flops_forward(rep, G, None)
    ops.flops_forward(rep, G, 150, want_stats=True)
    S = ops.scores(q.requires_grad_(True), rep.requires_grad_(True), True)
    loss = ops.rank_loss(S, None, "infonce", G, True)
    loss.backward()
    ops.compact_rows(rep)
    # encoder-body kernels on the step's activation shape [B*L, H]
    x = hidden.reshape(-1, H).clone().requires_grad_(True)
    gamma = torch.ones(H, device="cuda", requires_grad=True)
    beta = torch.zeros(H, device="cuda", requires_grad=True)
    y = ops.layer_norm(x, gamma, beta, 1e-12)
    y.backward(torch.ones_like(y))
    ops.colsum(hidden.reshape(-1, H))
    ops.colsum(torch.randn(B * L, 4 * H, device="cuda").bfloat16())
    # fused block tail and GELU kernels on the packed activation shape (85 % of B*L rows)
    T = int(B * L * 0.85) // 8 * 8
    yb = torch.randn(T, H, device="cuda").bfloat16().requires_grad_(True)
    res = torch.randn(T, H, device="cuda", requires_grad=True)
    for _ in range(2):
        o32, o16 = ops.add_layer_norm(yb, res, gamma, beta, 1e-12, p=0.1, training=True)
        torch.autograd.backward([o32, o16], [torch.randn_like(o32), torch.randn_like(o16)])
    pre = torch.randn(T, 4 * H, device="cuda").bfloat16()
    for _ in range(2):
        ops.gelu_forward(pre)
        ops.gelu_backward(pre, torch.randn_like(pre))
    torch.cuda.synchronize()

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "c2"
    shift = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    if which == "c2":
        run(160, 256, 384, 30522, 5, 32, shift)
    else:
        run(64, 512, 768, 30522, 2, 32, shift)
