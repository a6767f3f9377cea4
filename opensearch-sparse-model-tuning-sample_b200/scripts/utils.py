"""Hot-path helpers of the reference's ``scripts/utils.py``: ``gather_rep`` (NCCL all-gather whose local slice keeps
its gradient), ``is_ddp_enabled``, ``get_model``, ``set_logging``; plus ``DistEnv``, the small stand-in for the
``accelerate.Accelerator`` attributes the hot path uses (accelerate is not part of this stack).

The OpenSearch client / bulk / search helpers of the reference file are out of scope.
"""
import json
import logging
import os
import sys

import torch
import torch.distributed as dist


class _GatherKeepLocalGrad(torch.autograd.Function):
    """Rank-major all-gather; backward hands each rank the gradient slice of its own rows.

    Equivalent to the reference's ``all_rep = accelerator.gather(rep); all_rep[i*n:(i+1)*n] = rep``
    (scripts/utils.py:16-23): remote rows are constants, local rows carry grad.
    """

    @staticmethod
    def forward(ctx, rep, env):
        ctx.env = env
        ctx.rows = rep.shape[0]
        out = env.gather(rep.detach())
        # consumers that know about it (ops.FlopsFunction / ops.ScoresFunction) only compute the gradient rows that
        # survive this function's backward -- the local slice -- instead of all world_size * n rows
        lo = env.local_process_index * ctx.rows
        out._sb200_grad_rows = (lo, lo + ctx.rows)
        return out

    @staticmethod
    def backward(ctx, grad_all):
        lo = ctx.env.local_process_index * ctx.rows
        return grad_all[lo:lo + ctx.rows].contiguous(), None


def gather_rep(rep, accelerator):
    """``accelerator`` needs ``num_processes``, ``local_process_index`` and ``gather`` (DistEnv or accelerate). With a
    DistEnv whose peer sinks are enabled (trainer option rep_gather="peer") the exchange runs over symmetric NVLink peer
    memory instead of NCCL (scripts/peer.py) -- same rows, same gradient rule."""
    if accelerator.num_processes == 1:
        return rep
    sink = accelerator.peer_sink_for(rep) if hasattr(accelerator, "peer_sink_for") else None
    if sink is not None:
        from .peer import peer_gather
        return peer_gather(rep, sink)
    return _GatherKeepLocalGrad.apply(rep, accelerator)


class DistEnv:
    """One process per GPU over torch.distributed (NCCL on CUDA tensors, gloo on CPU tensors in tests)."""

    def __init__(self, group=None):
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.num_processes = dist.get_world_size(group) if on else 1
        self.process_index = dist.get_rank(group) if on else 0
        # single node: the reference indexes the gathered tensor with local_process_index (utils.py:21)
        self.local_process_index = self.process_index
        self.is_main_process = self.process_index == 0
        self.peer_sinks = None      # scripts.peer.PeerSinks once enable_peer_sinks() has been called
        self._peer_site = 0

    # ------------------------------------------------------------------ symmetric peer memory (rep_gather="peer")
    def enable_peer_sinks(self, device):
        """-> PeerSinks, or None (on every rank alike) when CUDA IPC / peer access is not available between the ranks:
        a one-element probe sink is created, written, signalled and read back before anything depends on it."""
        from .peer import PeerSinks, PeerUnavailable
        if self.peer_sinks is None:
            sinks = PeerSinks(self, device)
            try:
                probe = sinks.get("probe", 1, 4, torch.float32)
                self._in_step = True
                mine = torch.full((1, 4), float(self.process_index + 1), device=device)
                probe.fan_out(mine)
                probe.publish()
                probe.wait()
                want = torch.arange(1, self.num_processes + 1, device=device, dtype=torch.float32)
                good = bool((probe.gathered[:, 0] == want).all())
                flag = torch.tensor([1 if good else 0], dtype=torch.int32, device=device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
                if int(flag) == 0:
                    raise PeerUnavailable("peer probe read back wrong data")
            except PeerUnavailable as exc:
                logging.getLogger(__name__).warning("rep_gather='peer' is not available here, using NCCL: %s", exc)
                sinks.close()
                return None
            finally:
                self._in_step = False
            self.peer_sinks = sinks
        return self.peer_sinks

    def begin_step(self):
        """Called by the trainer at the start of a training step: peer gathers are allowed until end_step() (the reuse of
        a sink is ordered by the gradient all-reduce that follows every step)."""
        self._peer_site = 0
        self._in_step = True

    def end_step(self):
        self._in_step = False

    def peer_sink_for(self, tensor):
        """The sink a gather of `tensor` should use, or None (NCCL). A tensor the fused head already wrote into its sink is
        recognised by address; everything else gets the next numbered site of this step (sites are created collectively
        on first use, so every rank must issue the same gathers in the same order -- true for the training step)."""
        if self.peer_sinks is None or not getattr(self, "_in_step", False) or not tensor.is_cuda or tensor.dim() != 2:
            return None
        if (tensor.shape[0] * tensor.shape[1] * tensor.element_size()) % 16 != 0:
            return None
        for sink in self.peer_sinks._sinks.values():
            if sink.holds(tensor):
                return sink
        self._peer_site += 1
        return self.peer_sinks.get(f"site{self._peer_site}", tensor.shape[0], tensor.shape[1], tensor.dtype)

    def gather_plain(self, tensor):
        """Gather without autograd (ids, teacher outputs): peer memory when enabled, NCCL otherwise."""
        sink = self.peer_sink_for(tensor)
        if sink is None:
            return self.gather(tensor)
        from .peer import peer_gather
        return peer_gather(tensor, sink).detach()

    def gather(self, tensor):
        if self.num_processes == 1:
            return tensor
        tensor = tensor.contiguous()
        out = tensor.new_empty((self.num_processes * tensor.shape[0],) + tuple(tensor.shape[1:]))
        dist.all_gather_into_tensor(out, tensor, group=self.group)
        return out

    def wait_for_everyone(self):
        if self.num_processes > 1:
            dist.barrier(group=self.group)

    def unwrap_model(self, model):
        return model.module if hasattr(model, "module") else model


def is_ddp_enabled():
    return bool(dist.is_available() and dist.is_initialized())


def set_logging(training_args, log_file_name):
    level = getattr(logging, str(getattr(training_args, "log_level", "info")).upper(), logging.INFO)
    handlers = [logging.StreamHandler(sys.stdout)]
    if getattr(training_args, "output_dir", None):
        os.makedirs(training_args.output_dir, exist_ok=True)
        handlers.append(logging.FileHandler(os.path.join(training_args.output_dir, log_file_name)))
    logging.basicConfig(level=level, format="%(asctime)s - %(levelname)s - %(name)s - %(message)s",
                        datefmt="%m/%d/%Y %H:%M:%S", handlers=handlers, force=True)


def get_model(model_args, backbone=None, tokenizer=None):
    """Builds the SparseModel from ModelArguments; idf.json is read only for inf-free models (reference :50-68)."""
    from .model.sparse_encoders import SparseModel

    idf = None
    if model_args.inf_free and model_args.idf_path:
        with open(model_args.idf_path) as f:
            idf = json.load(f)
    return SparseModel(model_args.model_name_or_path, idf=idf, tokenizer_id=model_args.tokenizer_name,
                       idf_requires_grad=model_args.idf_requires_grad, prune_ratio=model_args.prune_ratio,
                       preprocess_func=model_args.preprocess_func, use_l0=model_args.use_l0, backbone=backbone,
                       tokenizer=tokenizer, fuse_body=getattr(model_args, "fuse_body", True),
                       unpad_capacity=getattr(model_args, "unpad_capacity", None),
                       attention=getattr(model_args, "attention", "auto"))
