"""Drop-in for the hot-path half of the reference's ``scripts/train/trainer.py``.

``ModelWrapper`` and ``SparseModelTrainer.{flops_value, get_lambda, compute_loss, _save, set_bi_encoder_teacher}``
keep the reference's names, arguments and arithmetic; the arithmetic runs in the sm_100a kernels (``ops``).
``transformers.Trainer`` cannot be used in this stack (accelerate is absent), so ``SparseModelTrainer`` carries its
own minimal data-parallel loop: one process per GPU, torch DDP (NCCL) for the gradient all-reduce, ``gather_rep``
for the representation all-gather, bf16/fp16 autocast around the model forward, AdamW + linear warm-up as built by
train_ir.py.
"""
import contextlib
import gc
import json
import logging
import os
import types

import torch

from ... import ops
from ..utils import DistEnv, gather_rep

logger = logging.getLogger(__name__)


class ModelWrapper(torch.nn.Module):
    """One forward for docs (always through the network) and queries (IDF lookup when inf_free) -- reference :18-49."""

    def __init__(self, sparse_model, inf_free=True):
        super().__init__()
        self.sparse_model = sparse_model
        self.inf_free = inf_free
        self.__dict__["peer_sinks"] = None   # scripts.peer.PeerSinks, set by a multi-GPU trainer (rep_gather="peer")

    def forward(self, inputs):
        sink = None
        sinks = self.__dict__.get("peer_sinks")
        if sinks is not None and self.training and self.sparse_model.prune_ratio is None:
            # the document vectors leave the head kernel straight into every rank's gathered buffer
            sink = sinks.get("d_rep", inputs["input_ids"].shape[0], self.sparse_model.vocab_size, torch.float32)
        d_rep = self.sparse_model(inf_free=False, _sink=sink, input_ids=inputs["input_ids"],
                                  attention_mask=inputs["attention_mask"])
        q_rep = self.sparse_model(inf_free=self.inf_free, input_ids=inputs["q_input_ids"],
                                  attention_mask=inputs["q_attention_mask"])
        return d_rep, q_rep

    def save(self, output_dir, **kwargs):
        sm = self.sparse_model
        sm.backbone.save_pretrained(output_dir, **kwargs)
        if sm.tokenizer is not None and hasattr(sm.tokenizer, "save_pretrained"):
            sm.tokenizer.save_pretrained(output_dir)
        if sm.idf_requires_grad:
            weights = sm.idf_vector.detach().cpu()
            table = {sm.tokenizer._convert_id_to_token(int(i)): float(weights[i]) for i in weights.nonzero().flatten()}
            with open(os.path.join(output_dir, "idf.json"), "w") as f:
                json.dump(table, f)


class SparseModelTrainer:
    def __init__(self, model_args, data_args, loss_functions, model=None, args=None, train_dataset=None,
                 data_collator=None, optimizers=(None, None), accelerator=None, grad_sync="ddp", rep_gather="nccl",
                 **unused):
        """grad_sync: "ddp" (torch DistributedDataParallel, bucketed all-reduce overlapped with backward; eager launches)
        or "flat" (gradients live in one flat fp32 buffer that is all-reduced with a single NCCL call after backward;
        this is the mode whose forward+backward can be captured in a CUDA graph on several GPUs), or "flat_overlap"
        (same buffer, cut into buckets whose all-reduces are issued from gradient hooks on a side stream while the
        backward pass is still running -- flat_grads.FlatGradBuckets; graph-capturable as well).
        rep_gather: "nccl" (all_gather_into_tensor, the reference's accelerate.gather) or "peer" (symmetric NVLink peer
        memory, scripts/peer.py: the head kernel stores the document vectors into every rank's gathered buffer from its
        epilogue; ids / scores / teacher vectors go through a copy kernel; flag barrier instead of a collective). "peer"
        applies to the training step on CUDA with 2..8 ranks of one node; everything else falls back to "nccl"."""
        self.model_args = model_args
        self.data_args = data_args
        self.loss_functions = loss_functions
        self.args = args
        self.train_dataset = train_dataset
        self.data_collator = data_collator
        self.optimizer, self.lr_scheduler = optimizers
        self.accelerator = accelerator if accelerator is not None else DistEnv()
        self.state = types.SimpleNamespace(global_step=0)
        self._ema = None
        self._ema_host = 0
        self.last_stats = None
        wrapper = ModelWrapper(model, model_args.inf_free)
        self.model_wrapper = wrapper
        self.model = wrapper
        self.grad_sync = grad_sync if self.accelerator.num_processes > 1 else "none"
        self.rep_gather = "nccl"
        if (rep_gather == "peer" and 1 < self.accelerator.num_processes <= 8 and model is not None
                and next(model.parameters()).is_cuda and hasattr(self.accelerator, "enable_peer_sinks")):
            wrapper.__dict__["peer_sinks"] = self.accelerator.enable_peer_sinks(next(model.parameters()).device)
            if wrapper.__dict__["peer_sinks"] is not None:      # None: no CUDA IPC / peer access here -> NCCL gather
                self.rep_gather = "peer"
        self._flat_grads = None
        self._buckets = None
        if self.grad_sync == "ddp" and next(wrapper.parameters()).is_cuda:
            dev = next(wrapper.parameters()).device
            self.model = torch.nn.parallel.DistributedDataParallel(wrapper, device_ids=[dev.index],
                                                                   gradient_as_bucket_view=True)
        elif self.grad_sync in ("flat", "flat_overlap"):
            from .flat_grads import FlatGradBuckets
            overlap = self.grad_sync == "flat_overlap"
            self._buckets = FlatGradBuckets(self.model_wrapper.parameters(), self.accelerator.num_processes,
                                            group=getattr(self.accelerator, "group", None), overlap=overlap,
                                            bucket_bytes=(24 << 20) if overlap else (1 << 62))
            self._flat_grads = self._buckets.flat
        self.scaler = None
        if args is not None and getattr(args, "fp16", False):
            self.scaler = torch.amp.GradScaler("cuda")
        # CUDA-graph mode (enable_cuda_graph): the whole step replays as one graph launch
        self._graph = None
        self._static_inputs = None
        self._static_loss = None
        self._step_t = None
        self._ovf_host = None
        self._ovf_event = None
        self._half_params = None

    # ------------------------------------------------------------------ reference attribute, without a per-step sync
    @property
    def ranking_loss_moving_avg(self):
        """0.99/0.01 EMA of the ranking loss (reference :120-122). Kept on the device; reading it synchronises."""
        return self._ema_host if self._ema is None else float(self._ema)

    @ranking_loss_moving_avg.setter
    def ranking_loss_moving_avg(self, value):
        self._ema, self._ema_host = None, value

    # ------------------------------------------------------------------ regularisers
    def flops_value(self, representation, group_num=1):
        """reference :61-73"""
        return ops.flops_value(representation, group_num, self.data_args.flops_threshold)

    def get_lambda(self, lambda_value, lambda_T):
        """reference :75-79. In CUDA-graph mode the same schedule is evaluated on the device from a step counter
        that lives inside the graph, so a replay never bakes in a stale host value."""
        if self._step_t is not None:
            ramp = torch.clamp((self._step_t + 1.0) / float(lambda_T), max=1.0)
            return float(lambda_value) * ramp * ramp
        if self.state.global_step >= lambda_T:
            return lambda_value
        return lambda_value * ((self.state.global_step + 1) / lambda_T) ** 2

    # ------------------------------------------------------------------ the hot loop body
    def compute_loss(self, model, inputs, return_outputs=False, num_items_in_batch=None):
        """reference :81-143"""
        if hasattr(self, "bi_encoder_teacher"):
            # HF Trainer runs compute_loss inside its autocast context, so the reference's teachers run in half precision
            with self._autocast():
                inputs["scores"] = self.bi_encoder_teacher.get_scores_batch(q_features_list=inputs["query"][1:],
                                                                            d_features_list=inputs["docs"][1:])
        student = {"q_input_ids": inputs["query"][0]["input_ids"],
                   "q_attention_mask": inputs["query"][0]["attention_mask"],
                   "input_ids": inputs["docs"][0]["input_ids"],
                   "attention_mask": inputs["docs"][0]["attention_mask"]}
        d_rep, q_rep = model(student)
        d_rep = gather_rep(d_rep, self.accelerator)
        q_rep = self._gather_queries(q_rep, student["q_input_ids"])
        if "scores" in inputs:
            inputs["scores"] = gather_rep(inputs["scores"], self.accelerator)

        d_flops = self.flops_value(d_rep, d_rep.shape[0] // q_rep.shape[0])
        flops_loss = d_flops * self.get_lambda(self.data_args.flops_d_lambda, self.data_args.flops_d_T)
        if not self.model_args.inf_free:
            flops_loss = flops_loss + self.flops_value(q_rep) * self.get_lambda(self.data_args.flops_q_lambda,
                                                                                 self.data_args.flops_q_T)
        ranking_loss = 0
        for loss_function in self.loss_functions:
            ranking_loss = ranking_loss + loss_function.get_loss(q_rep=q_rep, d_rep=d_rep, inputs=inputs)

        # moving average without the reference's per-step .item() host sync
        r = ranking_loss.detach().float()
        if self._ema is None:
            self._ema = torch.full_like(r, float(self._ema_host))
        self._ema.mul_(0.99).add_(r, alpha=0.01)

        loss = ranking_loss + flops_loss
        capturing = torch.cuda.is_current_stream_capturing() if d_rep.is_cuda else False
        if (self.args is not None and not capturing and self._graph is None
                and self.state.global_step % max(1, self.args.logging_steps) == 0):
            self._log_step(d_rep, d_flops, flops_loss)
        # DDP averages gradients over ranks while every rank holds the full global loss (reference :139-141)
        loss = loss * self.accelerator.num_processes
        return (loss, {"q_rep": q_rep, "d_rep": d_rep}) if return_outputs else loss

    def _gather_queries(self, q_rep, q_input_ids):
        """Cross-rank query vectors. Inf-free queries with a frozen IDF table are a pure function of the token ids, so
        the compact form is exchanged -- int64 ids [n_q, Lq] instead of dense fp32 [n_q, V] (~500x fewer bytes) -- and the
        vectors of all ranks are rebuilt locally by the IDF kernel (bit-exact, no gradient involved). Everything else
        goes through the reference's dense gather (scripts/utils.py:16-23)."""
        env = self.accelerator
        sm = self.model_wrapper.sparse_model
        if (env.num_processes > 1 and self.model_args.inf_free and not sm.idf_requires_grad and q_rep.is_cuda
                and hasattr(env, "gather")):
            # The collator pads to the longest text of the LOCAL batch (collator.py:34-41), so Lq differs between ranks
            # while all_gather_into_tensor needs equal shapes: every rank right-pads its ids to max_seq_length (the
            # collator truncates there, so no batch is longer) with a special-token id, which the IDF kernel ignores
            # exactly like the reference zeroes the special columns -- the rebuilt vectors stay bit-exact.
            width = int(getattr(self.data_args, "max_seq_length", 0) or 0)
            ids = q_input_ids
            if ids.shape[1] > width:
                width = self._agree_on_width(ids.shape[1])
            if ids.shape[1] < width:
                pad = ids.new_full((ids.shape[0], width - ids.shape[1]), int(sm.special_token_ids[0]))
                ids = torch.cat([ids, pad], dim=1)
            ids = ids.to(torch.int32).contiguous()
            all_ids = env.gather_plain(ids) if hasattr(env, "gather_plain") else env.gather(ids)
            return ops.idf_query(all_ids, sm.idf_vector, sm._special_ids_on(all_ids.device))
        return gather_rep(q_rep, env)

    def _agree_on_width(self, local_width):
        """max over ranks of a host integer (queries longer than max_seq_length: one small all-reduce + sync)."""
        import torch.distributed as dist
        dev = next(self.model_wrapper.parameters()).device
        t = torch.tensor([int(local_width)], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=getattr(self.accelerator, "group", None))
        return int(t.item())

    def _log_step(self, d_rep, d_flops, flops_loss):
        with torch.no_grad():
            _, _, _, stats = ops.flops_forward(d_rep.detach(), 1, None, want_stats=True)
            nnz, psum, pmax, _ = stats.tolist()
        self.last_stats = {"avg_doc_length": nnz / d_rep.shape[0], "nonzero_mean": psum / max(nnz, 1.0), "nonzero_max": pmax}
        logger.info("Step %d. ranking loss moving avg:%s, d_flops: %s, flops_loss: %s avg doc length: %s",
                    self.state.global_step, self.ranking_loss_moving_avg, float(d_flops.detach()), float(flops_loss.detach()),
                    self.last_stats["avg_doc_length"])
        logger.info("nonzero entries: %s %s %s", self.last_stats["nonzero_mean"], self.last_stats["nonzero_mean"],
                    self.last_stats["nonzero_max"])

    # ------------------------------------------------------------------ step / loop
    def _autocast(self):
        """fp16 / bf16 autocast exactly when the training arguments ask for it; fp32 otherwise, like the reference.
        (The fused head always multiplies half-precision operands: bf16 unless fp16 autocast is active.)"""
        if self.args is not None and getattr(self.args, "fp16", False):
            return torch.autocast("cuda", dtype=torch.float16)
        if self.args is None or getattr(self.args, "bf16", False):
            return torch.autocast("cuda", dtype=torch.bfloat16)
        return contextlib.nullcontext()

    def _zero_grads(self):
        if self._buckets is not None:
            self._buckets.prepare()     # gradients are written (not accumulated) and moved into the flat buffer per bucket
        else:
            self.optimizer.zero_grad(set_to_none=True)

    def _sync_flat_grads(self):
        """Moves the remaining gradients into the flat buffer and all-reduces them (mean over ranks, like DDP; the loss
        carries x world). "flat": one all-reduce; "flat_overlap": most buckets are already in flight."""
        self._buckets.finish()

    def _forward_backward(self, inputs):
        """forward (autocast) + loss + backward; gradients are left in .grad (not yet synchronised in "flat" mode)."""
        def run(student):
            with self._autocast():
                return self.model(student)

        self.model_wrapper.sparse_model.unpad_step_reset()
        self._refresh_half_weights()
        if hasattr(self.accelerator, "begin_step"):
            self.accelerator.begin_step()      # peer-memory gather sites are numbered per step
        loss = self.compute_loss(run, inputs)
        if self.scaler is not None:
            flag = self.model_wrapper.sparse_model.unpad_step_flag()
            if flag is not None:   # fp16: poison the loss so that GradScaler skips the step of an overflowed batch
                loss = loss + torch.where(flag > 0, torch.full_like(flag, float("inf")), torch.zeros_like(flag)).sum()
            self.scaler.scale(loss).backward()
        else:
            loss.backward()
        if hasattr(self.accelerator, "end_step"):
            self.accelerator.end_step()
        return loss

    def _refresh_half_weights(self):
        """One multi-tensor cast of the matrix parameters to the autocast dtype at the start of the step instead of one
        cast kernel per Linear per forward pass (ops.refresh_half_weights / ops.half_weight)."""
        if self.args is None or not next(self.model_wrapper.parameters()).is_cuda:
            return
        if getattr(self.args, "fp16", False):
            dtype = torch.float16
        elif getattr(self.args, "bf16", False):
            dtype = torch.bfloat16
        else:
            return
        if self._half_params is None:   # every matrix parameter + the biases of the Linear layers (not the LayerNorm affines)
            backbone = self.model_wrapper.sparse_model.backbone
            chosen = {id(p): p for p in backbone.parameters() if p.dim() == 2 and p.dtype == torch.float32}
            for m in backbone.modules():
                b = getattr(m, "bias", None)
                if isinstance(m, torch.nn.Linear) and b is not None and b.dtype == torch.float32:
                    chosen[id(b)] = b
            self._half_params = list(chosen.values())
        ops.refresh_half_weights(self._half_params, dtype)

    def _optimizer_step(self):
        """optimizer.step(), skipped ON THE DEVICE for a batch that overflowed the packed-body capacity: the fused
        AdamW kernels take the same `found_inf` flag GradScaler uses for fp16 overflows, so the weights, the moments
        and the step count stay untouched and a truncated batch can never leak into the model (it is counted, and
        check_unpad() raises)."""
        if self.scaler is not None:
            self.scaler.step(self.optimizer)
            self.scaler.update()
            return
        flag = self.model_wrapper.sparse_model.unpad_step_flag()
        if flag is not None:
            if not getattr(self.optimizer, "_step_supports_amp_scaling", False):
                raise RuntimeError("unpad_capacity < 1 needs an optimizer that can skip a step on a device flag "
                                   "(torch.optim.AdamW(..., fused=True)); use that or unpad_capacity=1.0")
            self.optimizer.grad_scale = None
            self.optimizer.found_inf = flag
        self.optimizer.step()

    def _eager_step_body(self, inputs):
        """forward + loss + backward + gradient sync + optimizer.step, no scheduler / bookkeeping."""
        loss = self._forward_backward(inputs)
        if self._flat_grads is not None:
            self._sync_flat_grads()
        if self.scaler is not None:
            self.scaler.unscale_(self.optimizer)
        max_norm = getattr(self.args, "max_grad_norm", None) if self.args is not None else None
        if max_norm:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm)
        self._optimizer_step()
        return loss

    @staticmethod
    def _copy_into(dst, src):
        if torch.is_tensor(dst):
            dst.copy_(src, non_blocking=True)
        elif hasattr(dst, "items"):
            for k in dst:
                SparseModelTrainer._copy_into(dst[k], src[k])
        else:
            for a, b in zip(dst, src):
                SparseModelTrainer._copy_into(a, b)

    def enable_cuda_graph(self, example_inputs, warmup_steps=3):
        """Captures the WHOLE step as one CUDA graph (fixed batch shapes): forward + loss + backward + optimizer, and on
        several GPUs also every collective -- the all-gathers of the representations, the gradient all-reduce(s)
        (bucketed and overlapped with backward in "flat_overlap" mode) -- so a step is a single graph launch.

        The PyTorch backbone issues ~1000 small launches per step and is host-bound in eager mode; replaying a graph
        removes that. Requirements: bf16 (no GradScaler), no gradient clipping, an optimizer built with
        capturable=True whose lr is a tensor, batches padded to the shapes of `example_inputs`, no live reference to a
        loss / autograd graph of an earlier eager step, and grad_sync "flat" / "flat_overlap" on several GPUs. The
        regulariser warm-up (get_lambda) is computed on the device from an in-graph step counter. Call
        release_graph() before tearing the process group down.
        """
        multi = self.accelerator.num_processes > 1
        if multi and self.grad_sync not in ("flat", "flat_overlap"):
            raise RuntimeError('CUDA-graph mode on several GPUs needs grad_sync="flat" or "flat_overlap" (gradients in '
                               "one flat buffer with static addresses)")
        if self.scaler is not None:
            raise RuntimeError("CUDA-graph mode supports bf16 only (fp16 needs GradScaler's host-side decisions)")
        if self.args is not None and getattr(self.args, "max_grad_norm", None):
            raise RuntimeError("CUDA-graph mode does not support gradient clipping")
        device = next(self.model_wrapper.parameters()).device
        self.model.train()
        self._static_inputs = self._to_device(example_inputs, device)
        self._static_inputs = {k: (list(v) if isinstance(v, list) else v) for k, v in self._static_inputs.items()}
        self._step_t = torch.full((), float(self.state.global_step), device=device, dtype=torch.float32)
        if self._ema is None:
            self._ema = torch.full((), float(self._ema_host), device=device, dtype=torch.float32)

        def whole_step():
            if self._flat_grads is not None:
                self._zero_grads()
            loss = self._eager_step_body(dict(self._static_inputs))
            self._step_t += 1.0
            return loss

        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(warmup_steps):
                if self._flat_grads is None:
                    self.optimizer.zero_grad(set_to_none=True)
                whole_step()
                if self.lr_scheduler is not None:
                    self.lr_scheduler.step()
                self.state.global_step += 1
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        graph = torch.cuda.CUDAGraph()
        if self._flat_grads is None:
            self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            self._static_loss = whole_step().detach()
        self._graph = graph
        # the capture itself does not execute; host-side counters stay where the warm-up left them
        return self

    def release_graph(self):
        """Drops the captured graph (and with it the captured NCCL work) so that the process group can be destroyed."""
        if self._graph is None:
            return
        dev = next(self.model_wrapper.parameters()).device
        torch.cuda.synchronize(dev)
        self._graph = None
        self._static_loss = None
        self._static_inputs = None
        self._step_t = None
        gc.collect()
        torch.cuda.synchronize(dev)

    def close(self):
        """Releases the captured graph and the peer-memory sinks (call before destroying the process group)."""
        self.release_graph()
        sinks = getattr(self.accelerator, "peer_sinks", None)
        if sinks is not None:
            sinks.close()
            self.accelerator.peer_sinks = None
            self.model_wrapper.__dict__["peer_sinks"] = None

    def training_step(self, inputs):
        """forward (autocast) + loss + backward + optimizer step; returns the detached loss tensor (no sync)."""
        self._poll_unpad()
        if self._graph is not None:
            self._copy_into(self._static_inputs, inputs)
            self._graph.replay()
            loss = self._static_loss
        else:
            self.model.train()
            if self._flat_grads is not None:
                self._zero_grads()
            loss = self._eager_step_body(inputs).detach()
            if self._flat_grads is None:
                self.optimizer.zero_grad(set_to_none=True)
        if self.lr_scheduler is not None:
            self.lr_scheduler.step()
        self.state.global_step += 1
        self._post_unpad()
        return loss

    def get_train_dataloader(self):
        if self.train_dataset is None:
            raise ValueError("Trainer: training requires a train_dataset.")
        from torch.utils.data import DataLoader
        from torch.utils.data.distributed import DistributedSampler
        sampler = None
        if self.accelerator.num_processes > 1 and not getattr(self.train_dataset, "no_prepare", False):
            sampler = DistributedSampler(self.train_dataset, num_replicas=self.accelerator.num_processes,
                                         rank=self.accelerator.process_index, shuffle=True, seed=self.args.seed)
        return DataLoader(self.train_dataset, batch_size=self.args.per_device_train_batch_size, sampler=sampler,
                          shuffle=sampler is None, collate_fn=self.data_collator,
                          num_workers=self.args.dataloader_num_workers, drop_last=self.args.dataloader_drop_last,
                          pin_memory=True)

    @staticmethod
    def _to_device(obj, device):
        if torch.is_tensor(obj):
            return obj.to(device, non_blocking=True)
        if hasattr(obj, "items"):   # dict or a tokenizer's BatchEncoding
            return {k: SparseModelTrainer._to_device(v, device) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return type(obj)(SparseModelTrainer._to_device(v, device) for v in obj)
        return obj

    def train(self):
        if self.args is not None and int(getattr(self.args, "gradient_accumulation_steps", 1) or 1) != 1:
            raise NotImplementedError("gradient_accumulation_steps != 1 is not supported by this trainer")
        device = next(self.model_wrapper.parameters()).device
        loader = self.get_train_dataloader()
        if device.type == "cuda":
            from ..dataset.collator import PrefetchLoader
            loader = PrefetchLoader(loader, device)   # pinned host batch -> H2D on a side stream, one step ahead
        epoch = 0
        while self.state.global_step < self.args.max_steps:
            if hasattr(loader.sampler, "set_epoch"):
                loader.sampler.set_epoch(epoch)
            for batch in loader:
                self.training_step(self._to_device(batch, device))
                if self.args.save_strategy == "steps" and self.state.global_step % self.args.save_steps == 0:
                    self._save(os.path.join(self.args.output_dir, f"checkpoint-{self.state.global_step}"))
                if self.state.global_step >= self.args.max_steps:
                    break
            epoch += 1
        self.check_unpad()
        return self.state.global_step

    # ------------------------------------------------------------------ packed-body overflow (unpad_capacity < 1)
    # A batch with more real tokens than the packed capacity cannot be seen on the host without a sync. It is handled
    # in three layers: (1) on the device, in the same step, the optimizer update of such a batch is skipped
    # (_optimizer_step), so the weights never see it; (2) after every step the overflow counter is copied to pinned host
    # memory asynchronously and the NEXT step raises as soon as that copy has landed; (3) check_unpad() synchronises and
    # raises -- called before every checkpoint and at the end of train().
    def _post_unpad(self):
        sm = self.model_wrapper.sparse_model
        counter = sm.unpad_counter()
        if counter is None:
            return
        if self._ovf_host is None:
            self._ovf_host = torch.zeros((), dtype=torch.int64).pin_memory()
            self._ovf_event = torch.cuda.Event()
        self._ovf_host.copy_(counter, non_blocking=True)
        self._ovf_event.record()

    def _poll_unpad(self):
        if self._ovf_event is not None and self._ovf_event.query() and int(self._ovf_host) > 0:
            self.check_unpad()

    def check_unpad(self):
        n = self.model_wrapper.sparse_model.unpad_overflows()
        if n:
            raise RuntimeError(f"{n} batch(es) held more real tokens than unpad_capacity allows; their optimizer updates "
                               "were skipped on the device (the weights are intact). Raise unpad_capacity (1.0 can never "
                               "overflow) and resume.")

    _check_unpad = check_unpad

    def _save(self, output_dir=None, state_dict=None):
        """reference :145-156 -- main process only, ModelWrapper.save layout."""
        self.check_unpad()
        output_dir = output_dir if output_dir is not None else self.args.output_dir
        os.makedirs(output_dir, exist_ok=True)
        logger.info("Saving model checkpoint to %s", output_dir)
        if self.accelerator.is_main_process:
            self.accelerator.unwrap_model(self.model).save(output_dir, state_dict=state_dict,
                                                           safe_serialization=getattr(self.args, "save_safetensors", True))

    def set_bi_encoder_teacher(self, embedding_service=None, models=None):
        """reference :158-178"""
        from .bi_encoder_wrapper import BiEncoderWrapper
        kw = self.data_args.kd_ensemble_teacher_kwargs
        self.bi_encoder_teacher = BiEncoderWrapper(types=kw["types"], model_ids=kw["model_ids"],
                                                   use_in_batch_negatives=self.data_args.use_in_batch_negatives,
                                                   score_scale=kw.get("score_scale", 30),
                                                   embedding_service=embedding_service, models=models)
        self.bi_encoder_teacher.accelerator = self.accelerator
        device = next(self.model_wrapper.parameters()).device
        for m in self.bi_encoder_teacher.models:
            m.to(device)
