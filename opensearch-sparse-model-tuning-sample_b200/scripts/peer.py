"""Symmetric peer memory for the data-parallel exchange steps (reference ``scripts/utils.py:16-23`` gather_rep =
``accelerator.gather`` + local write-back; ``bi_encoder_wrapper.py:130``; ``trainer.py:101-104``) without NCCL.

Every rank (one process per GPU, one node) allocates the same buffers through the C ABI (``sb200_peer_*``: cudaMalloc +
CUDA IPC), imports its peers' handles and then reads / writes every rank's copy from its own kernels over NVLink:

* ``PeerSink`` is one gather site: a ``[world * rows, width]`` buffer per rank plus a flag array and two device-side
  epoch counters. ``slot()`` is this rank's ``[rows, width]`` slice of its own buffer and ``remote_slots()`` the same
  slice inside every other rank's buffer -- the fused head kernel stores its ``rep`` rows into all of them while it is
  still multiplying (``ops.sparse_head(out=..., peer_ptrs=...)``). Tensors produced elsewhere are fanned out by the
  generic copy kernel. ``publish()`` + ``wait()`` are the cross-GPU barrier (release/acquire flags, kernels on the
  current stream: CUDA-graph capturable, no host synchronisation).
* ``peer_gather(rep, sink, env)`` is the autograd form of gather_rep on a sink: returns the gathered buffer; the
  gradient flows into the local rows only.

Reuse contract: a sink may be written again once every rank has consumed the previous contents. In the training step
the gradient all-reduce that separates two steps orders that (a rank starts step t+1 only after every rank has finished
the backward pass of step t, which contains its last read of the gathered tensors). Sinks are therefore enabled by the
trainer for its training step only; everything else uses the NCCL gather.
"""
import ctypes

import torch
import torch.distributed as dist

from .. import _lib

_CTRL_BYTES = 1024     # flags [<= 16] u32 at offset 0, send epoch at 256, wait epoch at 512


class _RawCuda:
    """Exposes a raw device pointer to torch through __cuda_array_interface__ (zero-copy, no ownership)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerUnavailable(RuntimeError):
    """Raised on EVERY rank when any rank could not allocate / export / import a sink (no CUDA IPC between the
    processes, no peer access between two devices): the caller then stays with the NCCL gather."""


class PeerSink:
    """One gather site. Collective constructor: every rank must create its sinks in the same order."""

    def __init__(self, env, rows, width, dtype, device):
        self.env = env
        self.world = env.num_processes
        self.rank = env.process_index
        self.rows, self.width, self.dtype = int(rows), int(width), dtype
        self.device = device
        self.item = torch.empty((), dtype=dtype).element_size()
        self.slot_bytes = self.rows * self.width * self.item
        if self.slot_bytes % 16 != 0:
            raise ValueError("PeerSink: rows * width * itemsize must be a multiple of 16 bytes")
        self.data_bytes = self.world * self.slot_bytes
        total = _CTRL_BYTES + self.data_bytes
        lib = _lib.load()
        local = ctypes.c_void_p()
        group = getattr(env, "group", None)
        self.ptrs, error = [], None
        with torch.cuda.device(device):
            handle = (ctypes.c_ubyte * 64)()
            try:
                _lib.check(lib.sb200_peer_alloc(total, ctypes.byref(local)), "sb200_peer_alloc")
                _lib.check(lib.sb200_peer_export(local, handle), "sb200_peer_export")
            except _lib.SparseB200Error as exc:
                error = exc
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
            everyone = torch.empty(self.world * 64, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(everyone, mine, group=group)
            handles = everyone.cpu().view(self.world, 64)
            for r in range(self.world):
                if error is not None:
                    break
                if r == self.rank:
                    self.ptrs.append(int(local.value))
                    continue
                raw = (ctypes.c_ubyte * 64)(*handles[r].tolist())
                p = ctypes.c_void_p()
                try:
                    _lib.check(lib.sb200_peer_import(raw, ctypes.byref(p)), "sb200_peer_import")
                    self.ptrs.append(int(p.value))
                except _lib.SparseB200Error as exc:
                    error = exc
            # every rank learns whether ALL ranks succeeded (also the barrier: everyone has imported before anyone writes)
            ok = torch.tensor([0 if error is not None else 1], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok) == 0:
                for r, p in enumerate(self.ptrs):
                    if r != self.rank:
                        lib.sb200_peer_close(ctypes.c_void_p(p))
                if local.value:
                    lib.sb200_peer_free(local)
                raise PeerUnavailable(f"symmetric peer memory unavailable (rank {self.rank}: {error})")
        self._local = int(local.value)
        self._holder = _RawCuda(self._local, total)
        self._bytes = torch.as_tensor(self._holder, device=device)
        self._c_ptrs = (ctypes.c_void_p * self.world)(*self.ptrs)
        self.gathered = self._bytes[_CTRL_BYTES:].view(dtype).view(self.world * self.rows, self.width)

    # ------------------------------------------------------------------ addresses
    def slot(self):
        """This rank's [rows, width] slice of its OWN gathered buffer."""
        return self.gathered[self.rank * self.rows:(self.rank + 1) * self.rows]

    def remote_slots(self):
        """Device pointers of this rank's slice inside every OTHER rank's gathered buffer."""
        off = _CTRL_BYTES + self.rank * self.slot_bytes
        return [p + off for r, p in enumerate(self.ptrs) if r != self.rank]

    # ------------------------------------------------------------------ kernels on the current stream
    def fan_out(self, tensor):
        """Copies `tensor` ([rows, width]) into this rank's slot of every rank's buffer (incl. its own)."""
        t = tensor.detach()
        if t.dtype != self.dtype or tuple(t.shape) != (self.rows, self.width):
            raise ValueError(f"PeerSink: expected {(self.rows, self.width)} {self.dtype}, got {tuple(t.shape)} {t.dtype}")
        t = t.contiguous()
        with torch.cuda.device(self.device):
            code = _lib.load().sb200_peer_allgather(t.data_ptr(), self.slot_bytes, self.rank, self.world, self._c_ptrs,
                                                    _CTRL_BYTES, torch.cuda.current_stream().cuda_stream)
        _lib.check(code, "sb200_peer_allgather")

    def publish(self):
        with torch.cuda.device(self.device):
            code = _lib.load().sb200_peer_signal(self._local + 256, self.rank, self.world, self._c_ptrs, 0,
                                                 torch.cuda.current_stream().cuda_stream)
        _lib.check(code, "sb200_peer_signal")

    def wait(self):
        with torch.cuda.device(self.device):
            code = _lib.load().sb200_peer_wait(self._local + 512, self._local, self.world,
                                               torch.cuda.current_stream().cuda_stream)
        _lib.check(code, "sb200_peer_wait")

    def holds(self, tensor):
        """True if `tensor` already IS this rank's slot (the producing kernel wrote it, and the remote copies, itself)."""
        return tensor.data_ptr() == self.slot().data_ptr() and tuple(tensor.shape) == (self.rows, self.width) \
            and tensor.dtype == self.dtype and tensor.is_contiguous()

    def close(self):
        lib = _lib.load()
        torch.cuda.synchronize(self.device)
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                lib.sb200_peer_close(ctypes.c_void_p(p))
        lib.sb200_peer_free(ctypes.c_void_p(self._local))
        self.ptrs = []


class _PeerGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rep, sink):
        if not sink.holds(rep):
            sink.fan_out(rep)
        sink.publish()
        sink.wait()
        out = sink.gathered
        lo = sink.rank * sink.rows
        ctx.rows = (lo, lo + sink.rows)
        # a fresh view object per call: consumers tag it (ops read _sb200_grad_rows) and autograd versions it
        out = out.view(sink.world * sink.rows, sink.width)
        out._sb200_grad_rows = ctx.rows
        return out

    @staticmethod
    def backward(ctx, grad_all):
        lo, hi = ctx.rows
        return grad_all[lo:hi].contiguous(), None


def peer_gather(rep, sink):
    """gather_rep (scripts/utils.py:16-23) through a PeerSink: rank-major rows of every rank; gradient only into the
    local rows."""
    return _PeerGather.apply(rep, sink)


class PeerSinks:
    """Named sinks of one trainer, created lazily (collectively) on first use with the shapes seen there."""

    def __init__(self, env, device):
        self.env, self.device = env, device
        self._sinks = {}

    def get(self, name, rows, width, dtype):
        key = (name, int(rows), int(width), dtype)
        if key not in self._sinks:
            self._sinks[key] = PeerSink(self.env, rows, width, dtype, self.device)
        return self._sinks[key]

    def close(self):
        for s in self._sinks.values():
            s.close()
        self._sinks = {}
