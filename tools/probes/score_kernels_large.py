"""Launches the score kernels at the 8-GPU global-batch shape (Nq 256 x Nd 2048) a few times: run under
`ncu --metrics gpu__time_duration.sum` to see the per-kernel durations behind bench.py's `extras` numbers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import sparse_b200
from sparse_b200 import ops
dev = torch.device("cuda")
V, nq, nd, lq = 30522, 256, 2048, 64
g = torch.Generator(device=dev).manual_seed(5)
d = torch.relu(torch.randn(nd, V, device=dev, generator=g))
ids = torch.randint(1000, V, (nq, lq), device=dev)
idf = torch.rand(V, device=dev)
sp = torch.tensor([0, 100, 101, 102, 103], dtype=torch.int32, device=dev)
q = ops.idf_query_forward(ids, idf, sp)
big = torch.empty(64 << 20, dtype=torch.float32, device=dev)
for _ in range(3):
    big.sum()
    ops.scores_forward(q, d, True)
    big.sum()
    ops.score_loss_forward(q, d, None, "infonce", nd // nq, True, q_nnz_bound=lq)
    big.sum()
    ops.flops_forward(d, 8, None)
    big.sum()
    ops.compact_rows(d)
torch.cuda.synchronize()
