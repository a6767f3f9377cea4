"""One forward + backward launch of the attention kernels at a body-layer shape (for ncu captures).
usage: python tools/probes/attn_once.py [c2|c3] [drop_p]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import sparse_b200  # noqa: E402,F401
from sparse_b200 import ops  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "c2"
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
nseq, lo, hi, h, d = (160, 128, 256, 12, 32) if shape == "c2" else (64, 256, 512, 12, 64)
g = torch.Generator().manual_seed(5)
lens = torch.randint(lo, hi + 1, (nseq,), generator=g).tolist()
T = sum(lens)
qkv = (torch.randn(T, 3, h, d, generator=g) * 1.5).to(torch.bfloat16).cuda()
dout = torch.randn(T, h, d, generator=g).to(torch.bfloat16).cuda()
cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32).cuda()
seed = torch.tensor([99], dtype=torch.int64).cuda() if p > 0 else None
for _ in range(2):
    out, lse = ops.attn_forward(qkv, cu, hi, 1 / math.sqrt(d), p, seed)
    dqkv = ops.attn_backward(qkv, out, dout, lse, cu, hi, 1 / math.sqrt(d), p, seed)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()), float(dqkv.float().abs().mean()))
