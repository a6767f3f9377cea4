"""Tensor-level entry points of the B200 neural-sparse hot path.

Each function validates its torch tensors, allocates outputs/workspace with torch (device memory plumbing only) and
calls the C ABI of libsparse_b200.so on the current CUDA stream. There is no CPU implementation here: CPU tensors are
rejected. The `*Function` classes wire the kernels into autograd.
"""
import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.SparseB200Error(
                "sparse_b200 ops run on CUDA tensors only (sm_100a kernels; there is no CPU fallback)")


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------------------- sparse head
def head_forward(hidden, weight, bias, attention_mask, use_l0=False, want_aux=True):
    """Fused MLM-decoder GEMM + mask + max-pool + log1p(relu) (sparse_encoders.py:108-114).

    hidden [B,L,H] bf16, weight [V,H] bf16, bias [V] fp32 or None, attention_mask [B,L] (int64/int32/uint8/bool).
    Returns (rep [B,V] fp32, xmax [B,V] fp32 | None, argmax [B,V] int32 | None).
    """
    _need_cuda(hidden, weight, bias, attention_mask)
    if hidden.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise TypeError("head_forward expects bf16 hidden and weight")
    B, L, H = hidden.shape
    V = weight.shape[0]
    if weight.shape[1] != H:
        raise ValueError("weight must be [V, H]")
    hidden = hidden.contiguous()
    weight = weight.contiguous()
    mask = attention_mask
    if mask.dtype == torch.bool:
        mask = mask.view(torch.uint8)
    if mask.dtype not in (torch.int64, torch.int32, torch.uint8):
        mask = (mask != 0).to(torch.uint8)
    mask = mask.contiguous()
    if tuple(mask.shape) != (B, L):
        raise ValueError("attention_mask must be [B, L]")
    if bias is not None:
        bias = bias.detach().float().contiguous()
    lib = _lib.load()
    dev = hidden.device
    rep = torch.empty(B, V, dtype=torch.float32, device=dev)
    xmax = torch.empty(B, V, dtype=torch.float32, device=dev) if want_aux else None
    argmax = torch.empty(B, V, dtype=torch.int32, device=dev) if want_aux else None
    ws_bytes = lib.sb200_head_fwd_workspace_bytes(B, L)
    ws = _workspace(ws_bytes, dev)
    with torch.cuda.device(dev):
        code = lib.sb200_head_fwd(_ptr(hidden), _ptr(weight), _ptr(bias), _ptr(mask), mask.element_size(), B, L, H, V,
                                  _lib.HEAD_L0 if use_l0 else 0, _ptr(rep), _ptr(xmax), _ptr(argmax), _ptr(ws),
                                  ws.numel(), _stream())
    _lib.check(code, "sb200_head_fwd")
    return rep, xmax, argmax
