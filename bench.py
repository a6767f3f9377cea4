#!/usr/bin/env python
"""Benchmark of the neural-sparse hot path (BASELINE.json metric: infoNCE train samples/sec, docs encoded/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5]
                    [--regime dense|trained]

Workloads (BASELINE.json `configs`):
  c2 (default, configs[1]) inf-free doc-v2-mini infoNCE step: 32 queries x (1 pos + 4 negs), seq 256
  c3 (configs[2])          BERT-base L0 kd step (config_l0): 32 queries x 2 docs, seq 512
  c4 (configs[3])          kd-ensemble: mini student, dense (24x1024) + sparse (BERT-base) teachers score the in-batch docs
  c5 (configs[4])          fused sparse-head microbench sweep vs the unfused PyTorch ops on the same GPU (no training step)

One "step" = synthetic token batch -> BERT body -> fused sparse head (tcgen05 kernel) -> IDF query kernel -> [NCCL
all-gather] -> FLOPS + score + loss kernels -> backward (sparse head scatter kernels, body) -> [gradient all-reduce]
-> AdamW. Rank 0 prints ONE JSON line. `--impl reference` times the same full-size step on the host CPU cores
(oracle/reference_runner.py: the reference's own modules when its tree is present, the oracle port otherwise).
"""
import argparse
import json
import math
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

V = 30522
WORKLOADS = {
    # configs[1]: inf-free doc-v2-mini infoNCE fine-tune step (1 pos + 4 negs, batch 32, seq 256)
    "c2": dict(name="inf-free doc-v2-mini infoNCE step: 32 queries x (1 pos + 4 negs), doc seq 256, query seq 32, V=30522",
               shape="mini", n_queries=32, docs_per_query=5, doc_len=256, query_len=32, loss="infonce", in_batch=True,
               use_l0=False, flops_threshold=None, flops_d_lambda=0.05, flops_d_T=200, teachers=[]),
    # configs[2]: BERT-base L0-enhanced (config_l0: kldiv, 2 docs/query, flops_threshold 150), 64 seq x 512 per GPU
    "c3": dict(name="BERT-base L0 kd step (config_l0): 32 queries x 2 docs, doc seq 512, query seq 32, V=30522",
               shape="base", n_queries=32, docs_per_query=2, doc_len=512, query_len=32, loss="kldiv", in_batch=False,
               use_l0=True, flops_threshold=150, flops_d_lambda=0.08, flops_d_T=40000, teachers=[]),
    # configs[3]: kd-ensemble (config_kd: kldiv, in-batch negatives, posnegs 1 pos + 2 negs, 12 queries per GPU, seq 512)
    "c4": dict(name="kd-ensemble step (config_kd): 12 queries x (1 pos + 2 negs), doc seq 512, query seq 32, mini student, "
                    "dense (24x1024) + sparse (BERT-base) teachers scoring the in-batch docs, V=30522",
               shape="mini", n_queries=12, docs_per_query=3, doc_len=512, query_len=32, loss="kldiv", in_batch=True,
               use_l0=False, flops_threshold=None, flops_d_lambda=0.002, flops_d_T=200,
               teachers=[("dense", "large"), ("sparse", "base")]),
}
HIDDEN = {"mini": 384, "base": 768, "tiny": 64, "large": 1024}
TRAINED_BIAS_SHIFT = {"mini": -3.3, "base": -3.6, "tiny": -1.0}  # decoder-bias shift for the "trained-like" regime
NCU_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "head_fwd_dram_traffic.json")


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


def shared_config(wl, args, world):
    """`config` of the JSON line: the workload only (identical for both arms; engine details live under `engine`)."""
    nq = wl["n_queries"]
    return {"workload": wl["name"], "regime": args.regime, "per_gpu_queries": nq,
            "per_gpu_docs": nq * wl["docs_per_query"], "doc_len": wl["doc_len"], "query_len": wl["query_len"], "vocab": V,
            "loss": wl["loss"], "in_batch_negatives": wl["in_batch"], "global_queries": world * nq,
            "parallelism": f"dp{world}",
            "backbone": f"random-init BertForMaskedLM {wl['shape']} (H={HIDDEN[wl['shape']]})",
            "l2": "no explicit flush: one step touches > 126 MB (activations, fp32 params, AdamW state)"}


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
def build_teachers(wl, device):
    """Random-init stand-ins of the kd-ensemble teachers (no checkpoints offline): dense = BERT body of the gte-large
    scale with CLS pooling, sparse = BertForMaskedLM base through the fused head (BiSparseModel)."""
    if not wl["teachers"]:
        return None
    import transformers
    from sparse_b200.scripts import synthetic
    from sparse_b200.scripts.train.bi_encoder_wrapper import BiSparseModel, DenseModel
    models = []
    for i, (kind, shape) in enumerate(wl["teachers"]):
        torch.manual_seed(100 + i)
        cfg = transformers.BertConfig(vocab_size=V, max_position_embeddings=512, **synthetic.MODEL_SHAPES[shape])
        if kind == "dense":
            models.append(DenseModel(None, backbone=transformers.BertModel(cfg, add_pooling_layer=False)))
        else:
            models.append(BiSparseModel(None, backbone=transformers.BertForMaskedLM(cfg),
                                        tokenizer=synthetic.SyntheticTokenizer(V)))
    return [m.to(device) for m in models]


def build_trainer(wl, args, device, accelerator=None, model=None, grad_sync=None, optimizer=True, rep_gather="nccl"):
    import sparse_b200  # noqa: F401
    from sparse_b200.scripts import synthetic
    from sparse_b200.scripts.args import DataTrainingArguments, ModelArguments, TrainingArguments
    from sparse_b200.scripts.train.loss import LOSS_CLS_MAP
    from sparse_b200.scripts.train.trainer import SparseModelTrainer

    if model is None:
        shift = TRAINED_BIAS_SHIFT[wl["shape"]] if args.regime == "trained" else 0.0
        model = synthetic.build_sparse_model(wl["shape"], idf_vector=idf_vector(), use_l0=wl["use_l0"], bias_shift=shift,
                                             fuse_body=not args.no_fused_body,
                                             unpad_capacity=args.unpad_capacity if args.unpad_capacity > 0 else None,
                                             attention=getattr(args, "attention", "auto"))
        model.to(device)
    model_args = ModelArguments(inf_free=True, use_l0=wl["use_l0"])
    teacher_kw = {}
    if wl["teachers"]:
        teacher_kw = {"types": [k for k, _ in wl["teachers"]], "model_ids": [s for _, s in wl["teachers"]], "score_scale": 30}
    data_args = DataTrainingArguments(loss_types=[wl["loss"]], use_in_batch_negatives=wl["in_batch"],
                                      flops_d_lambda=wl["flops_d_lambda"], flops_d_T=wl["flops_d_T"],
                                      flops_threshold=wl["flops_threshold"], sample_num_one_query=wl["docs_per_query"],
                                      kd_ensemble_teacher_kwargs=teacher_kw)
    targs = TrainingArguments(bf16=True, learning_rate=2e-5, weight_decay=0.01, warmup_steps=200, max_steps=2000,
                              logging_steps=10 ** 9, max_grad_norm=None,
                              per_device_train_batch_size=wl["n_queries"])
    losses = [LOSS_CLS_MAP[wl["loss"]](use_in_batch_negatives=wl["in_batch"], weight=1, temperature=1.0)]
    opt = sched = None
    if optimizer:
        # capturable + tensor lr: the optimizer step can live inside the CUDA graph and the scheduler updates lr in place
        opt = torch.optim.AdamW(model.parameters(), lr=torch.tensor(targs.learning_rate, device=device),
                                weight_decay=targs.weight_decay, fused=True, capturable=True)
        sched = torch.optim.lr_scheduler.LambdaLR(
            opt, lambda s: min(1.0, (s + 1) / targs.warmup_steps) * max(0.0, (targs.max_steps - s) / targs.max_steps))
    return SparseModelTrainer(model_args, data_args, losses, model=model, args=targs, optimizers=(opt, sched),
                              accelerator=accelerator, grad_sync=grad_sync or "ddp", rep_gather=rep_gather)


def host_batch(wl, rank, step):
    from sparse_b200.scripts import synthetic
    own_scores = wl["loss"] != "infonce" and not wl["teachers"]   # kd data carries the teacher scores of its own docs
    b = synthetic.train_batch(wl["n_queries"], wl["docs_per_query"], wl["doc_len"], wl["query_len"],
                              seed=1234 + 1000 * rank + step, with_scores=wl["docs_per_query"] if own_scores else None,
                              n_feature_sets=1 + len(wl["teachers"]))

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


def clone_inputs(b):
    # compute_loss adds keys ("scores" gather) to the dict it receives: hand it a shallow copy
    return {k: (list(v) if isinstance(v, list) else v) for k, v in b.items()}


def dist_parity(trainer, wl, args, device, batch, world, rank):
    """Driver-visible evidence for SURVEY 8(a6)/(e): the N-rank step (gather_rep, loss x world, mean-reduced gradients;
    reference utils.py:16-23, trainer.py:101-104,139-141) against ONE process running the concatenated global batch on
    rank 0. Dropout off (eval mode); same weights (they are replicated)."""
    import types
    import torch.distributed as dist
    sm = trainer.model_wrapper.sparse_model
    was_training = trainer.model.training
    trainer.model.eval()
    for t in getattr(getattr(trainer, "bi_encoder_teacher", None), "models", []):
        t.eval()
    out = {}
    try:
        flat = trainer._flat_grads
        params = [p for p in sm.parameters() if p.requires_grad]

        def grads():
            return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).detach().float().flatten()
                              for p in params])
        trainer._zero_grads()
        loss_dp = trainer._forward_backward(clone_inputs(batch))
        g_dp = None
        if flat is not None:
            trainer._sync_flat_grads()
            g_dp = grads()

        def gather_tree(o):
            if torch.is_tensor(o):
                parts = [torch.empty_like(o) for _ in range(world)]
                dist.all_gather(parts, o.contiguous())
                return torch.cat(parts, 0)
            if isinstance(o, dict):
                return {k: gather_tree(v) for k, v in o.items()}
            return [gather_tree(v) for v in o]
        global_batch = gather_tree({k: v for k, v in batch.items()})
        if rank == 0:
            solo_env = types.SimpleNamespace(num_processes=1, process_index=0, local_process_index=0, is_main_process=True,
                                             gather=lambda t: t, unwrap_model=lambda m: m)
            solo = build_trainer(wl, args, device, accelerator=solo_env, model=sm, grad_sync="none", optimizer=False)
            if hasattr(trainer, "bi_encoder_teacher"):
                solo.set_bi_encoder_teacher(models=trainer.bi_encoder_teacher.models)
            solo.model.eval()
            solo.state.global_step = trainer.state.global_step     # same regulariser warm-up weight (get_lambda)
            trainer._zero_grads()
            if trainer._buckets is not None:
                trainer._buckets.enabled = False     # this backward pass is rank 0's alone: no bucket all-reduces
            loss_1 = solo._forward_backward(global_batch)
            if trainer._buckets is not None:
                trainer._buckets.enabled = True
            a, b = float(loss_dp.detach()) / world, float(loss_1.detach())
            out = {"loss_dp_over_world": a, "loss_single_process_global_batch": b,
                   "loss_rel_err": abs(a - b) / max(abs(b), 1e-12), "tolerance": 1e-4}
            if g_dp is not None:
                g1 = grads()
                num = float((g_dp - g1).norm())
                den = float(g1.norm())
                out["grad_rel_l2_err"] = num / max(den, 1e-30)
                out["grad_note"] = ("flat fp32 gradient after the N-rank all-reduce(AVG) vs the single-process global-batch "
                                    "gradient; bf16 operands + atomic accumulation order differ, tolerance 2e-2")
            out["ok"] = bool(out["loss_rel_err"] <= 1e-4 and out.get("grad_rel_l2_err", 0.0) <= 2e-2)
        dist.barrier()
    finally:
        trainer._zero_grads()
        trainer.model.train(was_training)
    return out


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

    if args.workload == "c5":
        if rank == 0:
            emit(c5_line(args, device, load_peaks()))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    wl = WORKLOADS[args.workload]
    peaks = load_peaks()
    grad_sync = args.grad_sync if args.grad_sync != "auto" else ("flat_overlap" if args.graph else "ddp")
    rep_gather = args.rep_gather if args.rep_gather != "auto" else ("peer" if 1 < world <= 8 else "nccl")
    trainer = build_trainer(wl, args, device, grad_sync=grad_sync, rep_gather=rep_gather)
    teachers = build_teachers(wl, device)
    if teachers is not None:
        trainer.set_bi_encoder_teacher(models=teachers)
    n_pool = 4
    hosts = [host_batch(wl, rank, i) for i in range(n_pool)]
    resident = [trainer._to_device(h, device) for h in hosts]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- warm-up (eager)
    n_warm = max(3, args.warmup)
    for i in range(n_warm):
        trainer.training_step(clone_inputs(resident[i % n_pool]))
    barrier()

    parity = None
    if world > 1 and not args.no_dist_parity:
        parity = dist_parity(trainer, wl, args, device, resident[0], world, rank)
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

    # ---------------- CUDA-graph capture of the whole step (forward, loss, backward, gradient all-reduce, optimizer)
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

    def region_resident():
        barrier()
        e0.record()
        for i in range(args.steps):
            trainer.training_step(resident[i % n_pool] if graphed else clone_inputs(resident[i % n_pool]))
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    last = [None]

    def region_e2e():
        barrier()
        e0.record()
        for i in range(args.steps):
            if graphed:
                loss_t = trainer.training_step(hosts[i % n_pool])      # pinned host -> static device buffers -> replay
            else:
                loss_t = trainer.training_step(trainer._to_device(hosts[i % n_pool], device))
            last[0] = float(loss_t)                                    # device->host read of the step's loss
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    def max_over_ranks(x):
        t = torch.tensor([x], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---------------- timed regions. A = inputs resident in HBM ("value"), B = pinned host batch -> H2D -> step -> loss
    # read-back every step ("e2e"). Each region times EXACTLY --steps steps (events + barrier, max over ranks); the
    # pair is repeated (A, B, A, B, ...) until >= ~2 s have been timed per kind and the MEDIAN region is reported, so
    # that a 20-step region (0.2 s) is not at the mercy of one scheduling hiccup.
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_a = [max_over_ranks(region_resident())]
    ms_b = [max_over_ranks(region_e2e())]
    repeats = args.repeats if args.repeats > 0 else int(min(15, max(1, math.ceil(2000.0 / max(ms_a[0], 1e-3)))))
    for _ in range(repeats - 1):
        ms_a.append(max_over_ranks(region_resident()))
        ms_b.append(max_over_ranks(region_e2e()))
    clocks = sampler.stop() if sampler is not None else None
    ms_resident, ms_e2e = statistics.median(ms_a), statistics.median(ms_b)
    trainer.check_unpad()   # raises if any batch overflowed the packed capacity (such steps were skipped on the device)

    if rank == 0:
        nq = wl["n_queries"]
        nd = nq * wl["docs_per_query"]
        H = HIDDEN[wl["shape"]]
        value = world * nq * args.steps / (ms_resident / 1e3)
        e2e = world * nq * args.steps / (ms_e2e / 1e3)
        head_flops = 2.0 * nd * wl["doc_len"] * H * V          # algorithmic: every B*L position, padding included
        # flops the kernel really executes: real tokens rounded up to 16 per sequence (padding is skipped)
        lens = [b["docs"][0]["attention_mask"].sum(1) for b in hosts]
        toks = sum(float(((l + 15) // 16 * 16).sum()) for l in lens) / len(lens)
        exec_flops = 2.0 * toks * H * V
        fwd_ms = kernel_ms.get("head_fwd", [])
        bwd_ms = kernel_ms.get("head_bwd", [])
        n_student = len(fwd_ms) // max(1, n_eager)  # teachers call the head too (kd-ensemble): keep the student's calls
        if n_student > 1:
            fwd_ms = fwd_ms[n_student - 1::n_student]
        fwd_avg = sum(fwd_ms) / len(fwd_ms) if fwd_ms else float("nan")
        achieved = head_flops / (fwd_avg / 1e3) / 1e12
        # the per-launch time comes from the eager region, where the GPU idles between launches and boosts to its
        # maximum clock: the matching denominator is the burst cuBLAS figure (kernel timed alone), not the sustained one
        peak = peaks["bf16_tflops"]
        traffic, traffic_src = ncu_traffic(wl)
        sm = trainer.model_wrapper.sparse_model
        packed = sm.__dict__.get("_packed") is not None
        line = {
            "metric": "infonce_train_samples_per_sec", "value": round(value, 2), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": n_warm, "ms_per_step": round(ms_resident / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": shared_config(wl, args, world),
            "engine": {"body": (f"padding-free packed body: cuBLAS GEMMs (library); varlen attention "
                                f"({'this repo' if getattr(sm._packed, 'attention', 'auto') != 'flash' else 'flash_attn library'}), "
                                "embeddings, block tails (dropout+residual+LayerNorm), GELU, bias gradients on this "
                                "repo's sm_100a kernels; the fused head reads the packed [T,H] rows directly" if packed else
                                f"padded transformers body, {sm.fused_layers} LayerNorm/Linear modules on this repo's kernels"),
                       "unpad_capacity": args.unpad_capacity if packed else None,
                       "launch": ("CUDA graph replay (whole step incl. NCCL collectives and optimizer)" if graphed
                                  else "eager launches"),
                       "grad_sync": trainer.grad_sync, "rep_gather": getattr(trainer, "rep_gather", "dense")},
            "timing": {"repeats": repeats, "stat": "median over repeats of regions of exactly `steps` steps; CUDA events, "
                                                   "barrier + synchronize on both sides, max over ranks per region",
                       "region_ms": [round(x, 3) for x in ms_a], "e2e_region_ms": [round(x, 3) for x in ms_b]},
            "e2e": {"value": round(e2e, 2), "unit": "samples/s", "h2d_bytes_per_step": batch_bytes(hosts[0]),
                    "d2h_bytes_per_step": 4},
            "gpu_launches": int(round(launches_per_step * args.steps)),
            "gpu_launches_note": f"{launches_per_step:.0f} launches of this repo's kernels per step (counted in the eager "
                                 "region); in the timed regions they are nodes of the replayed CUDA graph",
            "clocks": clocks,
            "eager": {"ms_per_step": round(ms_eager, 4), "clocks": clocks_eager},
            "roofline": {"kernel": "head_fwd_kernel (fused vocab GEMM + mask + max-pool + log1p)", "bound": "tensor",
                         "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops burst ({peaks['source']}); sustained figure "
                                        f"{peaks['bf16_tflops_sustained']}",
                         "avg_launch_ms": round(fwd_avg, 4), "flops_per_launch": head_flops,
                         "executed_flops_per_launch": exec_flops,
                         "achieved_executed_only": round(exec_flops / (fwd_avg / 1e3) / 1e12, 1),
                         "frac_executed_only": round(exec_flops / (fwd_avg / 1e3) / 1e12 / peak, 4),
                         "note": "CUDA events on the launch stream around sb200_head_fwd inside eagerly launched "
                                 "training steps of the same workload; `achieved` uses the algorithmic 2*B*L*H*V flop "
                                 "of SURVEY.md 8(d) (all positions), `achieved_executed_only` counts only the token "
                                 "columns the kernel multiplies (real tokens, rounded up to 16 per sequence)"},
            "head_bwd_ms": round(sum(bwd_ms) / len(bwd_ms), 4) if bwd_ms else None,
            "last_loss": last[0],
        }
        if parity is not None:
            line["dist_parity"] = parity
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(wl, args, seconds=args.cpu_seconds)
        if not args.no_extras and world == 1:
            line["extras"] = extras(trainer, wl, args, device, peaks, hosts)
        emit(line)
    if world > 1:
        barrier()
        trainer.release_graph()   # the captured NCCL work must be gone before the communicator is torn down
        barrier()
        trainer.close()           # peer-memory sinks (CUDA IPC mappings)
        barrier()
        dist.destroy_process_group()


def ncu_traffic(wl):
    """dram__bytes_read.sum + dram__bytes_write.sum of one head_fwd launch at this workload's shape, read from the
    committed summary of an `ncu --set full` capture (profiles/head_fwd_dram_traffic.json, written by
    tools/ncu_summary.py); None when no capture of this shape has been committed."""
    try:
        with open(NCU_TRAFFIC_FILE) as f:
            table = json.load(f)
    except (OSError, ValueError):
        return None, "no capture file"
    key = f"{wl['shape']}_L{wl['doc_len']}_B{wl['n_queries'] * wl['docs_per_query']}"
    ent = table.get(key)
    if not ent:
        return None, f"no capture for {key} in {os.path.relpath(NCU_TRAFFIC_FILE, ROOT)}"
    return ent["dram_bytes"], ent.get("source", os.path.relpath(NCU_TRAFFIC_FILE, ROOT))


# ------------------------------------------------------------------------------------------------ extras (1 GPU)
def _time_cuda(fn, iters, flush=None, read_flush=False):
    """Mean CUDA-event time of fn() in ms. L2 is flushed before every call: by WRITING a buffer larger than L2 (the
    profiling recipe's method; it leaves the L2 full of dirty lines whose write-back then competes with the kernel's own
    reads) or, read_flush=True, by READING one (L2 full of clean lines)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total = 0.0
    for _ in range(iters):
        if flush is not None:
            if read_flush:
                flush.view(torch.float32).sum()
            else:
                flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters


def extras(trainer, wl, args, device, peaks, hosts):
    """Secondary numbers: doc-encode throughput (the metric's other half), the HBM-bound kernels against the copy peak,
    the unfused PyTorch training step on this GPU, and the C5 fused-head sweep."""
    from sparse_b200 import ops
    from sparse_b200.scripts import synthetic
    out = {}
    model = trainer.model_wrapper.sparse_model
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2
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
    # before every launch, once by writing 256 MB (`ms`, the recipe's method) and once by reading 256 MB
    # (`ms_clean_l2`); bytes are the algorithmic single-pass figures of DESIGN.md section 4.3.
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
            ms_clean = _time_cuda(fn, 10, flush, read_flush=True)
            gbs, gbs_clean = nbytes / (ms / 1e3) / 1e9, nbytes / (ms_clean / 1e3) / 1e9
            res[name] = {"ms": round(ms, 4), "GB/s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm, 4),
                         "ms_clean_l2": round(ms_clean, 4), "GB/s_clean_l2": round(gbs_clean, 1),
                         "frac_of_hbm_peak_clean_l2": round(gbs_clean / hbm, 4), "bytes": nbytes}
        # yardstick: torch's own read-only reduction over the same bytes, same two flush methods
        add("torch_sum_read_only_yardstick", lambda: d_rep.sum(), nd * V * 4)
        add("flops_fwd", lambda: ops.flops_forward(d_rep, G, None), nd * V * 4)
        add("flops_fwd_l0_threshold", lambda: ops.flops_forward(d_rep, G, 150), 2 * nd * V * 4)
        add("scores_fwd_in_batch", lambda: ops.scores_forward(q_rep, d_rep, True, q_nnz_bound=lq), (nq + nd) * V * 4)
        add("scores_fwd_in_batch_dense_queries", lambda: ops.scores_forward(q_dense, d_rep, True), (nq + nd) * V * 4)
        # scores + infoNCE loss + dS in one call (2 launches: query lists, persistent cooperative row kernel)
        add("score_loss_fwd_infonce_in_batch",
            lambda: ops.score_loss_forward(q_rep, d_rep, None, "infonce", nd // nq, True, q_nnz_bound=lq), (nq + nd) * V * 4)
        add("idf_query", lambda: ops.idf_query_forward(ids, model.idf_vector, sp), nq * lq * 12 + nq * V * 4)
        add("compact_rows", lambda: ops.compact_rows(d_rep), 2 * nd * V * 4)
        out[tag] = res
    nq, nd = wl["n_queries"], wl["n_queries"] * wl["docs_per_query"]
    suite("hbm_kernels_step_shape", nq, nd, wl["docs_per_query"], wl["query_len"], wl["flops_threshold"])
    suite("hbm_kernels_large_shape_nq256_nd2048", 256, 2048, 8, 64, 150)
    del flush
    torch.cuda.empty_cache()
    if not args.no_unfused:
        out["unfused_torch_gpu_step"] = unfused_gpu_step(wl, args, device, hosts)
    if not args.no_sweep:
        out["c5_sweep"] = c5_sweep(device, peaks, quick=True)
    return out


def unfused_gpu_step(wl, args, device, hosts, steps=10):
    """The same training step with the reference's unfused PyTorch ops on this GPU (bf16 autocast, stock transformers
    BertForMaskedLM incl. its decoder GEMM -> [B,L,V] logits, torch losses, fused torch AdamW): the same-box comparator
    named by SURVEY.md 2.2 / 8(d). Inputs resident in HBM; CUDA events; no kernels of this repository involved."""
    from baseline import unfused_torch as U
    from sparse_b200.scripts import synthetic
    if wl["teachers"]:
        return {"skipped": "kd-ensemble teachers are not part of the unfused single-model comparison"}
    backbone = synthetic.build_backbone(wl["shape"]).to(device)
    if args.regime == "trained":
        with torch.no_grad():
            backbone.cls.predictions.decoder.bias.add_(TRAINED_BIAS_SHIFT[wl["shape"]])
    step = U.UnfusedStep(backbone, idf_vector().to(device), [100, 102, 0, 101, 103], wl)
    resident = [{k: ([{kk: vv.to(device) for kk, vv in f.items()} for f in v] if isinstance(v, list) else v.to(device))
                 for k, v in h.items()} for h in hosts]
    for i in range(3):
        step(resident[i % len(resident)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(resident[i % len(resident)])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    peak_gb = torch.cuda.max_memory_allocated(device) / 2 ** 30
    del step, backbone, resident
    torch.cuda.empty_cache()
    return {"ms_per_step": round(ms, 3), "samples_per_sec": round(wl["n_queries"] / (ms / 1e3), 1), "steps": steps,
            "peak_mem_gib_process": round(peak_gb, 2),
            "what": "reference op sequence (sparse_encoders.py:107-114, loss.py, trainer.py:61-143) in eager PyTorch, bf16 "
                    "autocast, padded stock transformers body, fused torch AdamW; baseline/unfused_torch.py"}


# ------------------------------------------------------------------------------------------------ C5 sweep
def c5_sweep(device, peaks, quick=False, regimes=("dense",)):
    """BASELINE configs[4]: the fused sparse head (this repo's tcgen05 kernel + sparse backward) against the reference's
    unfused PyTorch ops (F.linear -> * mask -> max(dim=1) -> log1p(relu), sparse_encoders.py:108-114, and their
    autograd) on the same GPU under bf16 autocast. hidden ~ N(0,1) bf16 [B,L,H], W ~ N(0,0.02) (fp32 parameter, cast to
    bf16 by both sides), lengths uniform in [L/2, L], seed 7. forward = training forward (arg-max kept /
    autograd graph built); fwd+bwd adds the backward to hidden, W and bias. L2 flushed before every timed call."""
    from baseline import unfused_torch as U
    from sparse_b200 import ops
    rows = []
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    iters = 3 if quick else 5
    Bs = (16, 64, 160, 256)
    for regime in regimes:
        for H in (384, 768):
            for L in (128, 256, 512):
                for B in Bs:
                    g = torch.Generator(device=device).manual_seed(7)
                    hidden = torch.randn(B, L, H, device=device, generator=g).bfloat16().requires_grad_(True)
                    W = (torch.randn(V, H, device=device, generator=g) * 0.02).requires_grad_(True)
                    sigma = 0.02 * math.sqrt(H)
                    shift = 0.0
                    if regime == "trained":  # ~256 active vocabulary entries per document
                        p_tok = 256.0 / (V * 0.75 * L)
                        shift = -sigma * float(torch.special.ndtri(torch.tensor(1.0 - p_tok, dtype=torch.float64)))
                    bias = torch.full((V,), shift, device=device).requires_grad_(True)
                    lens = torch.randint(L // 2, L + 1, (B,), device=device, generator=g)
                    mask = (torch.arange(L, device=device)[None, :] < lens[:, None]).long()
                    g_out = torch.rand(B, V, device=device, generator=g)

                    def fused_fwd():
                        return ops.sparse_head(hidden, W, bias, mask, use_l0=False)

                    def unfused_fwd():
                        with torch.autocast("cuda", dtype=torch.bfloat16):
                            return U.sparse_head(hidden, W, bias, mask, use_l0=False)

                    def fb(fn):
                        def run():
                            hidden.grad = W.grad = bias.grad = None
                            fn().backward(g_out)
                        return run
                    res = {}
                    for name, fn in (("fused_fwd", fused_fwd), ("unfused_fwd", unfused_fwd), ("fused_fwd_bwd", fb(fused_fwd)),
                                     ("unfused_fwd_bwd", fb(unfused_fwd))):
                        for _ in range(2):
                            fn()
                        res[name] = _time_cuda(fn, iters, flush)
                    with torch.no_grad():
                        a = fused_fwd()
                        b = unfused_fwd().float()
                        active = float((a > 0).float().sum(1).mean())
                        # the unfused side rounds logits to bf16 before the max: compare at bf16 resolution
                        max_abs = float((a - b).abs().max())
                    flops = 2.0 * B * L * H * V
                    tf = flops / (res["fused_fwd"] / 1e3) / 1e12
                    rows.append({"regime": regime, "H": H, "L": L, "B": B,
                                 "fused_fwd_ms": round(res["fused_fwd"], 4), "unfused_fwd_ms": round(res["unfused_fwd"], 4),
                                 "speedup_fwd": round(res["unfused_fwd"] / res["fused_fwd"], 2),
                                 "fused_fwd_bwd_ms": round(res["fused_fwd_bwd"], 4),
                                 "unfused_fwd_bwd_ms": round(res["unfused_fwd_bwd"], 4),
                                 "speedup_fwd_bwd": round(res["unfused_fwd_bwd"] / res["fused_fwd_bwd"], 2),
                                 "fused_fwd_tflops": round(tf, 1), "frac_of_bf16_peak": round(tf / peaks["bf16_tflops"], 4),
                                 "docs_per_sec_fwd": round(B / (res["fused_fwd"] / 1e3), 1),
                                 "active_per_doc": round(active, 1), "max_abs_diff_vs_unfused_bf16": round(max_abs, 5)})
                    del hidden, W, bias, mask, g_out
                    torch.cuda.empty_cache()
    return rows


def c5_line(args, device, peaks):
    regimes = ("dense", "trained") if args.regime == "trained" else ("dense",)
    sampler = ClockSampler(device.index or 0)
    rows = c5_sweep(device, peaks, quick=False, regimes=regimes)
    clocks = sampler.stop()
    head = next(r for r in rows if (r["H"], r["L"], r["B"], r["regime"]) == (384, 256, 160, "dense"))
    return {"metric": "sparse_docs_encoded_per_sec", "value": head["docs_per_sec_fwd"], "unit": "docs/s", "n_gpus": 1,
            "steps": 5, "warmup": 2, "ms_per_step": head["fused_fwd_ms"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "C5 fused sparse-head microbench sweep: seq 128-512 x batch 16-256 x H 384/768, V=30522, "
                                   "vs the unfused PyTorch ops on the same GPU; headline row H=384 L=256 B=160",
                       "l2": "flushed (256 MB write) before every timed call"},
            "clocks": clocks, "peak_tflops": peaks["bf16_tflops"], "c5_sweep": rows}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_baseline(wl, args, seconds=20.0):
    """The full-size step on the host cores (reported beside the GPU number, not a target): 1 warm-up + as many timed
    steps as fit in `seconds` (at least 1)."""
    from oracle import reference_runner
    shift = TRAINED_BIAS_SHIFT[wl["shape"]] if args.regime == "trained" else 0.0
    step, kind, how = reference_runner.make_cpu_step(wl, shift, idf_vector())
    step()  # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        step()
        n += 1
        dt = time.perf_counter() - t0
        if dt > seconds or n >= 20:
            break
    return {"value": round(n * wl["n_queries"] / dt, 3), "unit": "samples/s", "cores": os.cpu_count(), "kind": kind,
            "sample": f"{n} full-size steps ({wl['n_queries']} queries x {wl['docs_per_query']} docs, seq {wl['doc_len']}) "
                      f"after 1 warm-up, fp32, torch CPU with {torch.get_num_threads()} threads; {how}"}


def run_reference(args):
    """The reference arm: the reference's CPU implementation of the same full-size step (same `config`) on the host
    cores. --steps / --warmup are honoured up to a wall-clock cap (--ref-seconds, default 400 s for the whole run): if
    the projected time exceeds it, first the warm-ups beyond one, then the timed steps are cut; the line states the
    numbers actually run."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    if args.workload == "c5":
        emit({"impl": "reference", "unavailable": "the C5 sweep carries its unfused PyTorch column itself"})
        return
    from oracle import reference_runner
    wl = WORKLOADS[args.workload]
    t_start = time.perf_counter()
    shift = TRAINED_BIAS_SHIFT[wl["shape"]] if args.regime == "trained" else 0.0
    step, kind, how = reference_runner.make_cpu_step(wl, shift, idf_vector())
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter() - t0                       # first (cold) step
    budget = args.ref_seconds - (time.perf_counter() - t_start)
    want_warm, want_steps = max(1, args.warmup), max(1, args.steps)
    fit = max(1, int(budget / max(t1, 1e-3)))           # steps that still fit (cold-step time is an upper bound)
    warm = want_warm
    if (warm - 1) + want_steps > fit:
        warm = 1
    steps = max(1, min(want_steps, fit - (warm - 1)))
    for _ in range(warm - 1):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    value = steps * wl["n_queries"] / dt
    sample = (f"{steps} timed full-size steps ({wl['n_queries']} queries x {wl['docs_per_query']} docs, seq {wl['doc_len']}) "
              f"after {warm} warm-up, fp32 torch CPU, {torch.get_num_threads()} threads; {how}; requested steps/warmup "
              f"{args.steps}/{args.warmup}, wall-clock cap {args.ref_seconds:.0f} s")
    emit({
        "impl": "reference", "metric": "infonce_train_samples_per_sec", "value": round(value, 3), "unit": "samples/s",
        "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": round(dt / steps * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": shared_config(wl, args, world),
        "cpu_baseline": {"value": round(value, 3), "unit": "samples/s", "cores": os.cpu_count(), "kind": kind,
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
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--repeats", type=int, default=0,
                    help="how many times the (resident, e2e) pair of exactly-`steps`-step regions is timed; the median "
                         "region is reported. 0 = as many as needed for >= 2 s per kind (at most 15)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c5"])
    ap.add_argument("--regime", default="dense", choices=["dense", "trained"],
                    help="dense = random-init decoder (about all 30522 columns active per doc); trained = decoder bias "
                         "shifted so that a few hundred columns are active, like a trained checkpoint")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch every kernel eagerly (torch DDP for the gradients). Default: the whole step -- on "
                         "several GPUs including the NCCL all-gathers, the bucketed gradient all-reduces overlapped with "
                         "backward, and the optimizer -- is replayed as one CUDA graph")
    ap.set_defaults(graph=True)
    ap.add_argument("--grad-sync", default="auto", choices=["auto", "ddp", "flat", "flat_overlap"],
                    help="gradient synchronisation on several GPUs: auto = flat_overlap (bucketed all-reduces issued during "
                         "backward) with the CUDA graph, ddp without; flat = one all-reduce after the backward pass")
    ap.add_argument("--rep-gather", default="auto", choices=["auto", "nccl", "peer"],
                    help="exchange of the representations between ranks: peer = symmetric NVLink peer memory (the head "
                         "kernel stores the document vectors into every rank's gathered buffer from its epilogue, flag "
                         "barrier instead of a collective); nccl = all_gather_into_tensor. auto = peer on 2..8 GPUs")
    ap.add_argument("--attention", default="auto", choices=["auto", "own", "flash"],
                    help="attention kernels of the padding-free body: own = csrc/attention.cu (head_dim 32 / 64), flash = the "
                         "flash_attn library's varlen kernel (A/B); auto = own where supported")
    ap.add_argument("--unpad-capacity", type=float, default=0.85,
                    help="padding-free encoder body: real tokens are packed into ceil(capacity * B * L) rows (the synthetic "
                         "lengths are uniform in [L/2, L], mean 0.75; an overflowing batch skips its optimizer step on the "
                         "device, is counted, and fails the run). 0 = padded body")
    ap.add_argument("--no-fused-body", action="store_true",
                    help="keep torch.nn.LayerNorm in the backbone (A/B of the fused LayerNorm kernels)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the C5 sweep in `extras`")
    ap.add_argument("--no-unfused", action="store_true", help="skip the unfused PyTorch GPU step in `extras`")
    ap.add_argument("--no-dist-parity", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-seconds", type=float, default=400.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
