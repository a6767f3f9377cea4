"""Fused score + ranking-loss call (ops.score_loss -> sb200_score_loss_fwd / sb200_scores_bwd) against the CPU oracle
(loss.py:25-43, 57-77, 86-107): every loss, both scoring modes, every gather mode of the row kernel (register lists,
streamed lists, dense fallback behind the device flag), vocabulary sizes that change the row partitioning and the
16-byte phase of the rows, the caller-promised non-zero bound, and local-row gradients.
Tolerance: loss rel 1e-4 / abs 1e-5, gradients rel 1e-3 / abs 1e-6 (fp32, different summation order)."""
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


def sparse_rows(n, V, nnz, gen, scale=1.0):
    out = torch.zeros(n, V)
    for i in range(n):
        k = nnz if isinstance(nnz, int) else int(nnz[i])
        cols = torch.randperm(V, generator=gen)[:k]
        out[i, cols] = (torch.rand(k, generator=gen) + 0.1) * scale
    return out


def run_case(ops, name, in_batch, Nq, G, V, q_nnz, T=1.0, bound=0, seed=0, d_density=0.05):
    g = torch.Generator().manual_seed(seed + Nq * 7 + G)
    q = sparse_rows(Nq, V, q_nnz, g, scale=0.5)
    d = torch.relu(torch.randn(Nq * G, V, generator=g)) * (torch.rand(Nq * G, V, generator=g) < d_density)
    C = Nq * G if in_batch else G
    teacher = None if name == "infonce" else torch.randn(Nq, C, generator=g) * 2
    qc, dc = q.cuda().requires_grad_(True), d.cuda().requires_grad_(True)
    if bound:
        qc._sb200_nnz_bound = bound
    loss = ops.score_loss(qc, dc, None if teacher is None else teacher.cuda(), name, G, in_batch, T)
    (loss * 1.7).backward()
    qr, dr = q.clone().requires_grad_(True), d.clone().requires_grad_(True)
    want = R.ranking_loss(name, qr, dr, teacher, in_batch, T)
    (want * 1.7).backward()
    torch.testing.assert_close(loss.detach().cpu(), want.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(qc.grad.cpu(), qr.grad, rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(dc.grad.cpu(), dr.grad, rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("name", ["infonce", "kldiv", "marginmse"])
@pytest.mark.parametrize("in_batch", [False, True])
def test_step_shape(ops, name, in_batch):
    # C2 shape: 32 inf-free queries (<= 32 tokens) x 160 docs, bound promised like IdfQueryFunction does
    run_case(ops, name, in_batch, 32, 5, 30522, 30, T=2.0, bound=32)


@pytest.mark.parametrize("Nq,G,V,nnz,bound", [
    (256, 2, 30522, 60, 64),      # register lists at the capacity of the block (tpq = 2)
    (256, 2, 30522, 100, 0),      # lists too long for registers: streamed from L2 (mode B), bound unknown
    (600, 1, 5000, 40, 64),       # more queries than threads/tpq allow in registers -> mode B
    (9, 4, 30522, 1000, 0),       # long learned-sparse lists
    (9, 4, 30522, 1500, 0),       # beyond the list capacity: dense fallback behind the device-side flag
    (5, 3, 70001, 50, 64),        # odd V (row phase changes per row), three parts per row
    (4, 2, 250002, 20, 32),       # multilingual-size vocabulary: many parts
    (3, 2, 17, 5, 0), (2, 1, 1, 1, 0), (1, 4, 300, 7, 8),
])
def test_in_batch_gather_modes(ops, Nq, G, V, nnz, bound):
    for name in ("infonce", "kldiv"):
        run_case(ops, name, True, Nq, G, V, min(nnz, V), bound=bound, d_density=0.2 if V < 1000 else 0.05)


def test_ragged_list_lengths_and_empty_queries(ops):
    g = torch.Generator().manual_seed(3)
    nnz = torch.randint(0, 200, (40,), generator=g)
    nnz[0] = 0
    nnz[7] = 0
    run_case(ops, "infonce", True, 40, 3, 30522, nnz, seed=5)
    run_case(ops, "marginmse", True, 40, 3, 30522, nnz, seed=6, T=0.5)


@pytest.mark.parametrize("Nq,G,V", [(32, 2, 30522), (64, 2, 30522), (300, 3, 999), (2, 40, 30522), (3, 1, 10)])
def test_own_docs_cluster_kernel(ops, Nq, G, V):
    for name in ("kldiv", "marginmse", "infonce"):
        if name == "marginmse" and G < 2:
            continue
        run_case(ops, name, False, Nq, G, V, min(50, V))


def test_scores_output_matches_plain_scores(ops):
    g = torch.Generator().manual_seed(11)
    q = sparse_rows(32, 30522, 32, g).cuda()
    d = (torch.relu(torch.randn(160, 30522, generator=g)) * (torch.rand(160, 30522, generator=g) < 0.05)).cuda()
    loss, S, dS, _ = ops.score_loss_forward(q, d, None, "infonce", 5, True, q_nnz_bound=32)
    # the fused call picks the single-launch gather kernel here, plain scores the streaming row kernel: same sums in a
    # different order
    torch.testing.assert_close(S, ops.scores_forward(q, d, True), rtol=1e-6, atol=1e-6)
    # without the promised bound the fused call streams rows as well: bit-identical to plain scores
    _, S_stream, _, _ = ops.score_loss_forward(q, d, None, "infonce", 5, True, q_nnz_bound=0)
    assert torch.equal(S_stream, ops.scores_forward(q, d, True))
    want = R.student_scores(q.cpu(), d.cpu(), True)
    torch.testing.assert_close(S.cpu(), want, rtol=1e-5, atol=1e-4)
    # deterministic: the same call twice gives bit-identical loss and gradient
    loss2, S2, dS2, _ = ops.score_loss_forward(q, d, None, "infonce", 5, True, q_nnz_bound=32)
    assert torch.equal(loss, loss2) and torch.equal(dS, dS2)


def test_idf_queries_carry_their_bound(ops):
    ids = torch.randint(1000, 30522, (8, 24)).cuda()
    idf = torch.rand(30522).cuda()
    q = ops.idf_query(ids, idf, torch.tensor([0, 100, 101, 102, 103], dtype=torch.int32, device="cuda"))
    assert q._sb200_nnz_bound == 24
    assert int((q != 0).sum(1).max()) <= 24


@pytest.mark.parametrize("rows,G,V,thr", [(160, 5, 30522, None), (160, 5, 30522, 150), (2048, 8, 30522, None),
                                           (2048, 8, 30522, 9000), (30, 1, 1001, None), (7, 7, 333, 2), (16, 2, 30522, 0)])
def test_flops_cluster_split_vs_oracle(ops, rows, G, V, thr):
    """Row-split cluster kernel of the FLOPS regulariser: every split factor, odd widths, L0 threshold."""
    g = torch.Generator().manual_seed(rows + V)
    rep = torch.relu(torch.randn(rows, V, generator=g) - 1.0)
    rc = rep.cuda().requires_grad_(True)
    val = ops.flops_value(rc, G, thr)
    (val * 0.3).backward()
    rr = rep.clone().requires_grad_(True)
    want = R.flops_value(rr, G, thr)
    (want * 0.3).backward()
    torch.testing.assert_close(val.detach().cpu(), want.detach(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(rc.grad.cpu(), rr.grad, rtol=1e-4, atol=1e-8)
    # deterministic value (fixed-order final sum)
    assert torch.equal(ops.flops_value(rep.cuda(), G, thr), ops.flops_value(rep.cuda(), G, thr))
