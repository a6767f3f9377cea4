"""world_size-2 gloo tests (CPU) of the data-parallel host logic: DistEnv.gather / gather_rep keep the reference's
rank-major layout and local-slice gradient (scripts/utils.py:16-23), and 'every rank holds the global loss, loss x
world, gradients averaged' reproduces the single-process global-batch gradient (trainer.py:139-141)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import reference_path as R


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sparse_b200  # noqa: F401
        from sparse_b200.scripts.utils import DistEnv, gather_rep, is_ddp_enabled
        assert is_ddp_enabled()
        env = DistEnv()
        assert env.num_processes == world and env.local_process_index == rank
        g = torch.Generator().manual_seed(100)
        nq, G, V = 3, 2, 23
        all_q = torch.relu(torch.randn(world * nq, V, generator=g))
        all_d = torch.relu(torch.randn(world * nq * G, V, generator=g))
        w = torch.randn(V, V, generator=g) * 0.3  # a shared "model": rep = relu(x @ w)
        wp = w.clone().requires_grad_(True)
        q_local = torch.relu(all_q[rank * nq:(rank + 1) * nq] @ wp)
        d_local = torch.relu(all_d[rank * nq * G:(rank + 1) * nq * G] @ wp)
        q_all = gather_rep(q_local, env)
        d_all = gather_rep(d_local, env)
        assert q_all.shape[0] == world * nq and d_all.shape[0] == world * nq * G
        # rank-major layout
        ref_q = torch.relu(all_q @ w)
        torch.testing.assert_close(q_all.detach(), ref_q)
        loss = (R.infonce_loss(q_all, d_all, True) + 0.1 * R.flops_value(d_all, G)) * world
        loss.backward()
        grad = wp.grad.clone()
        dist.all_reduce(grad)
        grad /= world  # DDP mean
        # single-process global batch
        ws = w.clone().requires_grad_(True)
        gl = R.infonce_loss(torch.relu(all_q @ ws), torch.relu(all_d @ ws), True) + 0.1 * R.flops_value(torch.relu(all_d @ ws), G)
        gl.backward()
        torch.testing.assert_close(grad, ws.grad, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(loss.detach() / world, gl.detach(), rtol=1e-6, atol=1e-6)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gather_rep_and_global_gradient_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_gather_rep_single_process_is_identity():
    import sparse_b200  # noqa: F401
    from sparse_b200.scripts.utils import DistEnv, gather_rep
    x = torch.randn(3, 5, requires_grad=True)
    assert gather_rep(x, DistEnv()) is x


def _bucket_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sparse_b200  # noqa: F401
        from sparse_b200.scripts.train.flat_grads import FlatGradBuckets

        def build():
            torch.manual_seed(7)
            emb = torch.nn.Embedding(50, 16)
            body = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 16))
            dec = torch.nn.Linear(16, 50)
            dec.weight = emb.weight                      # tied, like the MLM decoder
            unused = torch.nn.Linear(4, 4)               # never reached by the loss
            return torch.nn.ModuleList([emb, body, dec, unused])

        def loss_of(m, step):
            g = torch.Generator().manual_seed(1000 * step + rank)
            ids = torch.randint(0, 50, (6, 5), generator=g)
            return m[2](m[1](m[0](ids))).logsumexp(-1).mean() * world

        ref, net = build(), build()
        buckets = FlatGradBuckets(net.parameters(), world, bucket_bytes=2048)
        assert len(buckets.bounds) >= 3
        for step in range(3):
            # reference: plain backward, one all-reduce per parameter, mean
            for p in ref.parameters():
                p.grad = None
            loss_of(ref, step).backward()
            want = []
            for p in ref.parameters():
                g = torch.zeros_like(p) if p.grad is None else p.grad.clone()
                dist.all_reduce(g)
                want.append(g / world)
            buckets.zero()
            loss_of(net, step).backward()
            assert any(buckets._launched), "no bucket was reduced during backward"
            buckets.finish()
            assert all(buckets._launched)
            for p, w in zip(net.parameters(), want):
                assert p.grad.data_ptr() >= buckets.flat.data_ptr()          # still a view of the flat buffer
                torch.testing.assert_close(p.grad, w, rtol=1e-6, atol=1e-7)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_bucketed_flat_gradients_match_plain_allreduce_world2():
    """flat_grads.FlatGradBuckets: hooks reduce buckets during backward (tied weights, a parameter without gradient,
    several steps); result = mean over ranks of the plain gradients."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
