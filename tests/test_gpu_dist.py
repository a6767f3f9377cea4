"""2-GPU NCCL test of the data-parallel training path (skipped on a single-GPU box): every rank evaluates the global
loss from gathered reps (compact id exchange for inf-free queries, dense all-gather for docs), loss x world + DDP mean
must reproduce the single-process global-batch loss and parameter gradients."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(V, loss, in_batch, inf_free, single_process=False, grad_sync="ddp", capturable=False, rep_gather="nccl"):
    import sparse_b200  # noqa: F401
    from sparse_b200.scripts import synthetic
    from sparse_b200.scripts.args import DataTrainingArguments, ModelArguments, TrainingArguments
    from sparse_b200.scripts.train.loss import LOSS_CLS_MAP
    from sparse_b200.scripts.train.trainer import SparseModelTrainer
    from sparse_b200.scripts.utils import DistEnv
    env = DistEnv()
    if single_process:  # reference run on one rank: no DDP wrapper, no collectives
        env.num_processes, env.process_index, env.local_process_index = 1, 0, 0
    idf = torch.rand(V, generator=torch.Generator().manual_seed(3)) * 5
    model = synthetic.build_sparse_model("tiny", idf_vector=idf, vocab_size=V, seed=0, bias_shift=-0.1, dropout=0.0).cuda()
    margs = ModelArguments(inf_free=inf_free)
    dargs = DataTrainingArguments(loss_types=[loss], use_in_batch_negatives=in_batch, flops_d_lambda=0.05, flops_d_T=50,
                                  flops_q_lambda=0.02, flops_q_T=30)
    targs = TrainingArguments(bf16=True, logging_steps=10 ** 9, max_grad_norm=None)
    fns = [LOSS_CLS_MAP[loss](use_in_batch_negatives=in_batch, temperature=1.0)]
    opt = None
    if capturable:
        opt = torch.optim.AdamW(model.parameters(), lr=torch.tensor(1e-5, device="cuda"), fused=True, capturable=True)
    return SparseModelTrainer(margs, dargs, fns, model=model, args=targs, accelerator=env, grad_sync=grad_sync,
                              optimizers=(opt, None), rep_gather=rep_gather)


def _worker(rank, world, port, loss, in_batch, inf_free, out, grad_sync="ddp", rep_gather="nccl"):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from sparse_b200.scripts import synthetic
        V, nq, G = 1500, 4, 3
        flat = grad_sync in ("flat", "flat_overlap")
        tr = _build(V, loss, in_batch, inf_free, grad_sync=grad_sync, capturable=flat, rep_gather=rep_gather)
        assert tr.accelerator.num_processes == world and tr.grad_sync == grad_sync and tr.rep_gather == rep_gather
        # kd data carries the teacher scores of the query's own docs; with in-batch negatives the teachers would score
        # every gathered doc: [nq, world * nq * G] per rank
        n_scores = None if loss == "infonce" else (world * nq * G if in_batch else G)
        batches = [synthetic.train_batch(nq, G, 40, query_len=12, vocab_size=V, seed=70 + r, device="cuda",
                                         with_scores=n_scores) for r in range(world)]

        def run(student):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return tr.model(student)
        tr.accelerator.begin_step()   # peer-memory gather sites are numbered per step
        loss_v = tr.compute_loss(run, dict(batches[rank]))
        loss_v.backward()  # DDP averages the gradients
        tr.accelerator.end_step()
        if flat:
            tr._sync_flat_grads()  # all-reduce of the flat buffer (or of its remaining buckets), mean over ranks
        grads = {n: p.grad.detach().float().clone() for n, p in tr.model_wrapper.named_parameters() if p.grad is not None}
        if rank == 0:
            # single-process global batch with the same weights
            ref = _build(V, loss, in_batch, inf_free, single_process=True)
            ref.model_wrapper.load_state_dict(tr.model_wrapper.state_dict())

            def cat(key, idx):
                return torch.cat([b[key][0][idx] for b in batches], 0)
            gb = {"query": [{"input_ids": cat("query", "input_ids"), "attention_mask": cat("query", "attention_mask")}],
                  "docs": [{"input_ids": cat("docs", "input_ids"), "attention_mask": cat("docs", "attention_mask")}]}
            if loss != "infonce":
                gb["scores"] = torch.cat([b["scores"] for b in batches], 0)

            def run_ref(student):
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return ref.model(student)
            ref_loss = ref.compute_loss(run_ref, gb)
            ref_loss.backward()
            torch.testing.assert_close(loss_v.detach() / world, ref_loss.detach(), rtol=1e-4, atol=1e-5)
            # bf16 activations with different batch compositions: compare the whole gradient in L2 and every sizeable
            # parameter by direction
            num = den = 0.0
            big = max(float(p.grad.float().norm()) for _, p in ref.model_wrapper.named_parameters() if p.grad is not None)
            for n, p in ref.model_wrapper.named_parameters():
                if p.grad is None:
                    continue
                a, b = grads[n].flatten(), p.grad.float().flatten()
                num += float((a - b).pow(2).sum())
                den += float(b.pow(2).sum())
                if float(b.norm()) > 1e-2 * big:
                    cos = float(torch.nn.functional.cosine_similarity(a, b, dim=0))
                    assert cos > 0.995, (n, cos)
            assert (num / den) ** 0.5 < 3e-2, (num / den) ** 0.5
            ref_loss = None
        if flat:
            # CUDA-graph replay of the whole step (collectives captured) against eager steps of a twin trainer
            # drop every reference to the eager autograd graph first: a live AccumulateGrad node created on the default
            # stream would be reused inside the capture and invalidate it
            loss_v = None
            import gc
            gc.collect()
            twin = _build(V, loss, in_batch, inf_free, grad_sync="flat", capturable=True, rep_gather=rep_gather)
            twin.model_wrapper.load_state_dict(tr.model_wrapper.state_dict())
            tr.enable_cuda_graph(batches[rank], warmup_steps=2)
            for _ in range(2):
                twin.training_step(dict(batches[rank]))
            for _ in range(3):
                lg = float(tr.training_step(batches[rank]))
                le = float(twin.training_step(dict(batches[rank])))
                assert abs(lg - le) <= 2e-3 * abs(le) + 1e-4, (lg, le)
            tr.release_graph()   # the captured NCCL work is gone before the process group is torn down
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        out.put((rank, traceback.format_exc()[-1500:]))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("loss,in_batch,inf_free,grad_sync,rep_gather", [
    ("infonce", True, True, "ddp", "nccl"), ("kldiv", False, False, "ddp", "nccl"), ("infonce", True, True, "flat", "nccl"),
    ("infonce", True, True, "flat_overlap", "peer"),   # fused head all-gather + id exchange over NVLink peer memory
    ("kldiv", True, False, "flat", "peer"),            # learned queries: q_rep and the teacher scores through copy sinks
])
def test_two_gpu_global_loss_and_gradients(loss, in_batch, inf_free, grad_sync, rep_gather):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, loss, in_batch, inf_free, out, grad_sync, rep_gather))
             for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=180) for _ in procs]
    hung = []
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():     # a rank that does not exit after its result is a teardown hang
            hung.append(p.pid)
            p.kill()
    assert sorted(r for r, _ in results) == [0, 1]
    assert all(msg == "ok" for _, msg in results), "\n".join(f"rank {r}: {m}" for r, m in results)
    assert not hung, f"ranks did not exit after finishing: {hung}"
