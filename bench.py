#!/usr/bin/env python
"""Benchmark of the neural-sparse hot path (BASELINE.json metric: infoNCE train samples/sec, docs encoded/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3] [--regime dense|trained]

One "step" = one infoNCE fine-tune step of configs[1] (inf-free doc-v2-mini shape: 32 queries x (1 pos + 4 negs),
doc seq 256, query seq 32, vocab 30522): synthetic token batch -> BERT body (PyTorch, bf16 autocast) -> fused sparse
head (tcgen05 kernel) -> IDF query kernel -> [NCCL all-gather] -> FLOPS + in-batch score + infoNCE kernels ->
backward (sparse head scatter kernels, PyTorch body) -> [DDP all-reduce] -> AdamW.  Rank 0 prints ONE JSON line.

`--impl reference` times the same step on the host CPU cores with the CPU restatement of the reference path
(oracle/reference_path.py; the reference itself is plain PyTorch and is not present on the GPU box) on a bounded
sample of the workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # configs[1]: inf-free doc-v2-mini infoNCE fine-tune step (1 pos + 4 negs, batch 32, seq 256)
    "c2": dict(name="inf-free doc-v2-mini infoNCE step: 32 queries x (1 pos + 4 negs), doc seq 256, query seq 32, V=30522",
               shape="mini", n_queries=32, docs_per_query=5, doc_len=256, query_len=32, loss="infonce", in_batch=True,
               use_l0=False, flops_threshold=None, flops_d_lambda=0.05, flops_d_T=200),
    # configs[2]: BERT-base L0-enhanced (config_l0: kldiv, 2 docs/query, flops_threshold 150), 64 seq x 512 per GPU
    "c3": dict(name="BERT-base L0 kd step (config_l0): 32 queries x 2 docs, doc seq 512, query seq 32, V=30522",
               shape="base", n_queries=32, docs_per_query=2, doc_len=512, query_len=32, loss="kldiv", in_batch=False,
               use_l0=True, flops_threshold=150, flops_d_lambda=0.08, flops_d_T=40000),
}
TRAINED_BIAS_SHIFT = {"mini": -3.3, "base": -3.6, "tiny": -1.0}  # decoder-bias shift for the "trained-like" regime


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def idf_vector():
    return torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "idf_vector_f32.npy")))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ our arm
def build_trainer(wl, regime, device):
    import sparse_b200  # noqa: F401
    from sparse_b200.scripts import synthetic
    from sparse_b200.scripts.args import DataTrainingArguments, ModelArguments, TrainingArguments
    from sparse_b200.scripts.train.loss import LOSS_CLS_MAP
    from sparse_b200.scripts.train.trainer import SparseModelTrainer

    shift = TRAINED_BIAS_SHIFT[wl["shape"]] if regime == "trained" else 0.0
    model = synthetic.build_sparse_model(wl["shape"], idf_vector=idf_vector(), use_l0=wl["use_l0"], bias_shift=shift,
                                         fuse_body=not getattr(build_trainer, "no_fused_body", False),
                                         unpad_capacity=getattr(build_trainer, "unpad_capacity", None))
    model.to(device)
    model_args = ModelArguments(inf_free=True, use_l0=wl["use_l0"])
    data_args = DataTrainingArguments(loss_types=[wl["loss"]], use_in_batch_negatives=wl["in_batch"],
                                      flops_d_lambda=wl["flops_d_lambda"], flops_d_T=wl["flops_d_T"],
                                      flops_threshold=wl["flops_threshold"], sample_num_one_query=wl["docs_per_query"])
    targs = TrainingArguments(bf16=True, learning_rate=2e-5, weight_decay=0.01, warmup_steps=200, max_steps=2000,
                              logging_steps=10 ** 9, max_grad_norm=None,
                              per_device_train_batch_size=wl["n_queries"])
    losses = [LOSS_CLS_MAP[wl["loss"]](use_in_batch_negatives=wl["in_batch"], weight=1, temperature=1.0)]
    # capturable + tensor lr: the optimizer step can live inside the CUDA graph and the scheduler updates lr in place
    opt = torch.optim.AdamW(model.parameters(), lr=torch.tensor(targs.learning_rate, device=device),
                            weight_decay=targs.weight_decay, fused=True, capturable=True)
    sched = torch.optim.lr_scheduler.LambdaLR(
        opt, lambda s: min(1.0, (s + 1) / targs.warmup_steps) * max(0.0, (targs.max_steps - s) / targs.max_steps))
    return SparseModelTrainer(model_args, data_args, losses, model=model, args=targs, optimizers=(opt, sched),
                              grad_sync=getattr(build_trainer, "grad_sync", "ddp"))


def host_batch(wl, rank, step):
    from sparse_b200.scripts import synthetic
    n_scores = None
    if wl["loss"] != "infonce":
        n_scores = wl["docs_per_query"]  # kd data carries per-query teacher scores of its own docs
    b = synthetic.train_batch(wl["n_queries"], wl["docs_per_query"], wl["doc_len"], wl["query_len"],
                              seed=1234 + 1000 * rank + step, with_scores=n_scores)

    def pin(o):
        if torch.is_tensor(o):
            return o.pin_memory()
        if isinstance(o, dict):
            return {k: pin(v) for k, v in o.items()}
        return [pin(v) for v in o]
    return pin(b)


def batch_bytes(b):
    if torch.is_tensor(b):
        return b.numel() * b.element_size()
    if isinstance(b, dict):
        return sum(batch_bytes(v) for v in b.values())
    return sum(batch_bytes(v) for v in b)


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")  # required for NCCL inside CUDA graphs
        dist.init_process_group("nccl", device_id=device)
    import sparse_b200
    from sparse_b200 import ops
    lib = sparse_b200._lib

    wl = WORKLOADS[args.workload]
    peaks = load_peaks()
    build_trainer.no_fused_body = args.no_fused_body
    build_trainer.grad_sync = args.grad_sync if args.grad_sync != "auto" else ("flat" if args.graph else "ddp")
    build_trainer.unpad_capacity = args.unpad_capacity if args.unpad_capacity > 0 else None
    trainer = build_trainer(wl, args.regime, device)
    n_pool = 4
    hosts = [host_batch(wl, rank, i) for i in range(n_pool)]
    resident = [trainer._to_device(h, device) for h in hosts]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def clone_inputs(b):
        # compute_loss adds keys ("scores" gather) to the dict it receives: hand it a shallow copy
        return {k: (list(v) if isinstance(v, list) else v) for k, v in b.items()}

    # ---------------- warm-up (eager)
    for i in range(max(3, args.warmup)):
        trainer.training_step(clone_inputs(resident[i % n_pool]))
    barrier()

    # ---------------- region E: eager launches with CUDA events around the head kernels (per-launch durations,
    # launch count). Also the un-graphed step time, reported for reference.
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ops.start_event_profile(["head_fwd", "head_bwd"])
    launches0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_eager = min(args.steps, 10)
    barrier()
    e0.record()
    for i in range(n_eager):
        trainer.training_step(clone_inputs(resident[i % n_pool]))
    e1.record()
    barrier()
    ms_eager = e0.elapsed_time(e1) / n_eager
    launches_per_step = (lib.launch_count() - launches0) / n_eager
    kernel_ms = ops.stop_event_profile()
    clocks_eager = sampler.stop() if sampler is not None else None

    # ---------------- CUDA-graph capture of the whole step (forward, loss, backward, optimizer)
    graphed = False
    if args.graph:
        try:
            trainer.enable_cuda_graph(hosts[0], warmup_steps=3)
            graphed = True
        except Exception as exc:  # capture refused (e.g. driver / NCCL combination): keep launching eagerly
            print(f"[bench] CUDA-graph capture failed, staying eager: {exc!r}", file=sys.stderr, flush=True)
            trainer._graph = None
        for i in range(3):
            trainer.training_step(resident[i % n_pool] if graphed else clone_inputs(resident[i % n_pool]))
        barrier()

    # ---------------- timed region A: inputs resident in HBM ("value")
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    e0.record()
    for i in range(args.steps):
        trainer.training_step(resident[i % n_pool] if graphed else clone_inputs(resident[i % n_pool]))
    e1.record()
    barrier()
    ms_resident = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler is not None else None

    # ---------------- timed region B: host buffers, H2D copy + loss read-back inside the region ("e2e")
    barrier()
    e0.record()
    last = None
    for i in range(args.steps):
        if graphed:
            loss_t = trainer.training_step(hosts[i % n_pool])      # pinned host -> static device buffers -> replay
        else:
            loss_t = trainer.training_step(trainer._to_device(hosts[i % n_pool], device))
        last = float(loss_t)                                       # device->host read of the step's loss
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    t = torch.tensor([ms_resident, ms_e2e], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_resident, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        nq = wl["n_queries"]
        nd = nq * wl["docs_per_query"]
        H = {"mini": 384, "base": 768, "tiny": 64}[wl["shape"]]
        V = 30522
        value = world * nq * args.steps / (ms_resident / 1e3)
        e2e = world * nq * args.steps / (ms_e2e / 1e3)
        head_flops = 2.0 * nd * wl["doc_len"] * H * V          # algorithmic: every B*L position, padding included
        # the kernel skips the padded tail of each sequence in 16-token steps (one sequence per tile when L > 128):
        # flops it really executes, averaged over the batches of the pool
        if wl["doc_len"] > 128:
            lens = [b["docs"][0]["attention_mask"].sum(1) for b in hosts]
            toks = sum(float(((l + 15) // 16 * 16).sum()) for l in lens) / len(lens)
        else:
            toks = float(nd * wl["doc_len"])
        exec_flops = 2.0 * toks * H * V
        fwd_ms = kernel_ms.get("head_fwd", [])
        bwd_ms = kernel_ms.get("head_bwd", [])
        fwd_avg = sum(fwd_ms) / len(fwd_ms) if fwd_ms else float("nan")
        # the algorithmic figure counts every B*L position (2*L*H*V flop per document); the flops the
        # kernel really multiplies (padding skipped) are reported next to it
        achieved = head_flops / (fwd_avg / 1e3) / 1e12
        # the per-launch time comes from the eager region, where the GPU idles between launches and boosts to its
        # maximum clock: the matching denominator is the burst cuBLAS figure (kernel timed alone), not the sustained one
        peak = peaks["bf16_tflops"]
        stats = trainer.last_stats
        overflows = trainer.model_wrapper.sparse_model.unpad_overflows()
        if overflows:
            raise SystemExit(f"{overflows} batches did not fit --unpad-capacity {args.unpad_capacity}: numbers invalid")
        line = {
            "metric": "infonce_train_samples_per_sec", "value": round(value, 2), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms_resident / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl["name"], "regime": args.regime, "per_gpu_queries": nq, "per_gpu_docs": nd,
                       "global_queries": world * nq, "parallelism": f"dp{world}",
                       "backbone": f"random-init BertForMaskedLM {wl['shape']}, bf16 autocast; "
                                   + ("padding-free body: cuBLAS GEMMs + flash_attn varlen (library), embeddings / "
                                      "block tails (dropout+residual+LayerNorm) / GELU / bias gradients on this "
                                      "repo's sm_100a kernels" if trainer.model_wrapper.sparse_model.__dict__.get("_packed")
                                      else f"padded transformers body, {trainer.model_wrapper.sparse_model.fused_layers} "
                                           "LayerNorm/Linear modules on this repo's sm_100a kernels"),
                       "unpad_capacity": args.unpad_capacity if trainer.model_wrapper.sparse_model.__dict__.get("_packed") else None,
                       "l2": "no explicit flush: one step touches > 126 MB (activations, fp32 params, AdamW state)",
                       "launch": ("CUDA graph replay" + (" (fwd+bwd captured, flat grad all-reduce + optimizer after)"
                                                         if world > 1 else " (whole step)")) if graphed
                       else "eager launches", "grad_sync": trainer.grad_sync},
            "e2e": {"value": round(e2e, 2), "unit": "samples/s", "h2d_bytes_per_step": batch_bytes(hosts[0]),
                    "d2h_bytes_per_step": 4},
            "gpu_launches": int(round(launches_per_step * args.steps)),
            "gpu_launches_note": f"{launches_per_step:.0f} launches of this repo's kernels per step (counted in the eager "
                                 "region); in the timed regions they are nodes of the replayed CUDA graph",
            "clocks": clocks,
            "eager": {"ms_per_step": round(ms_eager, 4), "clocks": clocks_eager},
            "roofline": {"kernel": "head_fwd_kernel (fused vocab GEMM + mask + max-pool + log1p)", "bound": "tensor",
                         "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s",
                         "frac": round(achieved / peak, 4),
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this shape, from the committed
                         # `ncu --set full` capture (profiles/r01_ncu_full_body_kernels_c2.txt); the algorithmic minimum
                         # is hidden 31.5 MB + W 23.4 MB + three [B,V] outputs 58.6 MB (part of which is still in L2)
                         "traffic": 88.0e6 if (wl["shape"] == "mini" and wl["doc_len"] == 256) else None,
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops burst ({peaks['source']}); sustained figure "
                                        f"{peaks['bf16_tflops_sustained']}",
                         "avg_launch_ms": round(fwd_avg, 4), "flops_per_launch": head_flops,
                         "executed_flops_per_launch": exec_flops,
                         "achieved_executed_only": round(exec_flops / (fwd_avg / 1e3) / 1e12, 1),
                         "note": "CUDA events on the launch stream around sb200_head_fwd (mask-pack kernel + fused "
                                 "kernel) inside eagerly launched training steps of the same workload; `achieved` "
                                 "uses the algorithmic 2*B*L*H*V flop of SURVEY.md 8(d) (all positions), "
                                 "`achieved_executed_only` counts only the token columns the kernel multiplies (the "
                                 "padded tail of each sequence is skipped in 16-token steps)"},
            "head_bwd_ms": round(sum(bwd_ms) / len(bwd_ms), 4) if bwd_ms else None,
            "last_loss": last,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(wl, args.regime, seconds=args.cpu_seconds)
        if not args.no_extras and world == 1:
            line["extras"] = extras(trainer, wl, device, peaks)
        emit(line)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        if graphed:
            # destroy_process_group() blocks while a CUDA graph that captured NCCL work is alive; the line is out,
            # every rank is past the barrier: leave without the teardown
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ extras (1 GPU)
def _time_cuda(fn, iters, flush=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters


def extras(trainer, wl, device, peaks):
    """Secondary numbers: doc-encode throughput (the metric's other half) and HBM-bound kernels against the copy peak."""
    from sparse_b200 import ops
    from sparse_b200.scripts import synthetic
    out = {}
    model = trainer.model_wrapper.sparse_model
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2
    V = 30522
    # docs encoded / s: C1 shape (batch 8 x seq 128) and a throughput shape (batch 256 x seq 128), forward only
    for name, (B, L) in {"encode_b8_l128": (8, 128), "encode_b256_l128": (256, 128)}.items():
        feats = synthetic.token_batch(B, L, seed=7, device=device)
        model.eval()

        def enc():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                return model(inf_free=False, **feats)
        for _ in range(3):
            enc()
        ms = _time_cuda(enc, 10, flush)
        out[name + "_docs_per_sec"] = round(B / (ms / 1e3), 1)
    model.train()
    # HBM-bound kernels: at the step's shapes (launch-latency regime: a few MB per call) and at a large shape
    # (8-GPU global batch of BERT-base C3 scale) where the HBM roofline is the meaningful yardstick. L2 is flushed
    # before every launch; bytes are the algorithmic single-pass figures of DESIGN.md section 4.3.
    hbm = peaks["hbm_gbs"]
    sp = model._special_ids_on(device)

    def suite(tag, nq, nd, G, lq, thr):
        g = torch.Generator(device=device).manual_seed(5)
        d_rep = torch.relu(torch.randn(nd, V, device=device, generator=g))
        ids = synthetic.token_batch(nq, lq, seed=9, device=device)["input_ids"]
        q_rep = ops.idf_query_forward(ids, model.idf_vector, sp)       # inf-free queries: <= lq non-zeros per row
        q_dense = torch.relu(torch.randn(nq, V, device=device, generator=g))  # learned-query worst case (dense init)
        res = {}

        def add(name, fn, nbytes):
            for _ in range(3):
                fn()
            ms = _time_cuda(fn, 10, flush)
            gbs = nbytes / (ms / 1e3) / 1e9
            res[name] = {"ms": round(ms, 4), "GB/s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm, 4), "bytes": nbytes}
        add("flops_fwd", lambda: ops.flops_forward(d_rep, G, None), nd * V * 4)
        add("flops_fwd_l0_threshold", lambda: ops.flops_forward(d_rep, G, 150), 2 * nd * V * 4)
        add("scores_fwd_in_batch", lambda: ops.scores_forward(q_rep, d_rep, True), (nq + nd) * V * 4)
        add("scores_fwd_in_batch_dense_queries", lambda: ops.scores_forward(q_dense, d_rep, True), (nq + nd) * V * 4)
        add("idf_query", lambda: ops.idf_query_forward(ids, model.idf_vector, sp), nq * lq * 12 + nq * V * 4)
        add("compact_rows", lambda: ops.compact_rows(d_rep), 2 * nd * V * 4)
        out[tag] = res
    nq, nd = wl["n_queries"], wl["n_queries"] * wl["docs_per_query"]
    suite("hbm_kernels_step_shape", nq, nd, wl["docs_per_query"], wl["query_len"], wl["flops_threshold"])
    suite("hbm_kernels_large_shape_nq256_nd2048", 256, 2048, 8, 64, 150)
    return out


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_train_step_factory(wl, regime, n_queries):
    """The reference path restated on CPU (fp32, all host threads): BERT body + decoder (PyTorch), then the oracle's
    head / IDF / FLOPS / loss arithmetic, backward and AdamW."""
    from oracle import reference_path as R
    from sparse_b200.scripts import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    backbone = synthetic.build_backbone(wl["shape"])
    if regime == "trained":
        with torch.no_grad():
            backbone.cls.predictions.decoder.bias.add_(TRAINED_BIAS_SHIFT[wl["shape"]])
    idf = idf_vector()
    special = [100, 102, 0, 101, 103]
    opt = torch.optim.AdamW(backbone.parameters(), lr=2e-5, weight_decay=0.01)
    batch = synthetic.train_batch(n_queries, wl["docs_per_query"], wl["doc_len"], wl["query_len"], seed=99,
                                  with_scores=None if wl["loss"] == "infonce" else wl["docs_per_query"])
    docs, queries = batch["docs"][0], batch["query"][0]

    def step():
        logits = backbone(**docs)[0]
        values = torch.max(logits * docs["attention_mask"].unsqueeze(-1), dim=1).values
        d_rep = R.activation(values, wl["use_l0"])
        q_rep = R.idf_query(queries["input_ids"], idf, special)
        loss, _, _, _ = R.compute_loss(q_rep, d_rep, loss_specs=[dict(name=wl["loss"], use_in_batch_negatives=wl["in_batch"])],
                                       global_step=0, flops_d_lambda=wl["flops_d_lambda"], flops_d_T=wl["flops_d_T"],
                                       flops_threshold=wl["flops_threshold"], teacher_scores=batch.get("scores"))
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return float(loss.detach())
    return step


def cpu_baseline(wl, regime, seconds=20.0, n_queries=2):
    step = cpu_train_step_factory(wl, regime, n_queries)
    step()  # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        step()
        n += 1
        dt = time.perf_counter() - t0
        if dt > seconds or n >= 20:
            break
    return {"value": round(n * n_queries / dt, 3), "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{n} steps of {n_queries} queries x {wl['docs_per_query']} docs (seq {wl['doc_len']}), fp32, "
                      f"torch CPU with {torch.get_num_threads()} threads; oracle/reference_path.py arithmetic"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    n_queries = 2
    step = cpu_train_step_factory(wl, args.regime, n_queries)
    for _ in range(min(args.warmup, 1) or 1):
        step()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    value = steps * n_queries / dt
    sample = (f"each step = {n_queries} queries x {wl['docs_per_query']} docs (seq {wl['doc_len']}) of the workload, fp32 "
              f"torch CPU, {torch.get_num_threads()} threads")
    emit({
        "impl": "reference", "metric": "infonce_train_samples_per_sec", "value": round(value, 3), "unit": "samples/s",
        "n_gpus": world, "steps": steps, "warmup": 1, "ms_per_step": round(dt / steps * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "regime": args.regime, "parallelism": "cpu"},
        "cpu_baseline": {"value": round(value, 3), "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


_RESULT_OUT = None


def claim_stdout():
    """stdout carries exactly one JSON line. Libraries write banners to fd 1 (NCCL prints "NCCL version ..." there when
    NCCL_DEBUG is set on the box): keep a private duplicate of the original stdout for the result and point fd 1 at
    stderr for everything else."""
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    if os.environ.get("SB200_FAULT_TIMEOUT"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["SB200_FAULT_TIMEOUT"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--regime", default="dense", choices=["dense", "trained"],
                    help="dense = random-init decoder (about all 30522 columns active per doc); trained = decoder bias "
                         "shifted so that a few hundred columns are active, like a trained checkpoint")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch every kernel eagerly (torch DDP for the gradients). Default: the step is replayed as a "
                         "CUDA graph -- the whole step on one GPU; forward + backward (incl. the NCCL all-gathers) on "
                         "several GPUs, followed by one flat gradient all-reduce and the optimizer")
    ap.set_defaults(graph=True)
    ap.add_argument("--grad-sync", default="auto", choices=["auto", "ddp", "flat", "flat_overlap"],
                    help="gradient synchronisation on several GPUs: auto = flat (one all-reduce after the backward pass) "
                         "with the CUDA graph, ddp without; flat_overlap = bucketed all-reduces issued during backward")
    ap.add_argument("--unpad-capacity", type=float, default=0.85,
                    help="padding-free encoder body: real tokens are packed into ceil(capacity * B * L) rows (the synthetic "
                         "lengths are uniform in [L/2, L], mean 0.75; overflows are counted and fail the run). 0 = padded")
    ap.add_argument("--no-fused-body", action="store_true",
                    help="keep torch.nn.LayerNorm in the backbone (A/B of the fused LayerNorm kernels)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
