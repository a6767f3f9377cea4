import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_b200
from sparse_b200 import ops
B, L, H, V = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (160, 256, 384, 30522))]
FULL = "full" in sys.argv[5:]
NOAUX = "noaux" in sys.argv[5:]
g = torch.Generator(device="cuda").manual_seed(0)
hidden = torch.randn(B, L, H, device="cuda", generator=g).bfloat16()
W = (torch.randn(V, H, device="cuda", generator=g) * 0.05).bfloat16()
bias = torch.randn(V, device="cuda", generator=g) * 0.1
lens = torch.randint(L // 2, L + 1, (B,), device="cuda", generator=g)
if FULL:
    lens = torch.full((B,), L, device="cuda")
mask = (torch.arange(L, device="cuda")[None, :] < lens[:, None]).long()
for _ in range(3):
    ops.head_forward(hidden, W, bias, mask, want_aux=not NOAUX)
torch.cuda.synchronize()
