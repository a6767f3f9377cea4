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


# Optional per-call CUDA-event timing of named C-ABI calls (used by bench.py for the live roofline numbers).
_PROFILE = None


def start_event_profile(names):
    global _PROFILE
    _PROFILE = {n: [] for n in names}


def stop_event_profile():
    """-> {name: [milliseconds per call]}; synchronises the device."""
    global _PROFILE
    prof, _PROFILE = _PROFILE, None
    if not prof:
        return {}
    torch.cuda.synchronize()
    return {n: [a.elapsed_time(b) for a, b in pairs] for n, pairs in prof.items()}


class _timed:
    def __init__(self, name):
        self.pair = None
        if _PROFILE is not None and name in _PROFILE:
            self.pair = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            _PROFILE[name].append(self.pair)

    def __enter__(self):
        if self.pair is not None:
            self.pair[0].record()

    def __exit__(self, *exc):
        if self.pair is not None:
            self.pair[1].record()


# --------------------------------------------------------------------------------------------- half-precision weights
# Under autocast every Linear re-casts its fp32 master weight to bf16 / fp16 in every forward pass (53 small cast kernels
# per C2 step, the 23 MB decoder weight among them). refresh_half_weights() does all of them with one multi-tensor copy
# at the start of a step; the ops below then pick the cached copy up as long as the parameter has not been modified
# since (tensor version check), and fall back to a private cast otherwise.
def refresh_half_weights(params, dtype=torch.bfloat16):
    """params: list of fp32 parameters (any shapes). Creates (once) and refreshes their half-precision copies."""
    params = [p for p in params if p.is_cuda and p.dtype == torch.float32]
    if not params:
        return
    halves = []
    for p in params:
        c = p.__dict__.get("_sb200_half")
        if c is None or c[0].dtype != dtype or c[0].shape != p.shape:
            c = [torch.empty_like(p, dtype=dtype), -1]
            p.__dict__["_sb200_half"] = c
        halves.append(c[0])
    torch._foreach_copy_(halves, [p.detach() for p in params])
    for p in params:
        p.__dict__["_sb200_half"][1] = p._version


def half_weight(weight, dtype):
    """weight in `dtype`: the cached copy when it is current, a fresh cast otherwise (no-op when already `dtype`)."""
    if weight.dtype == dtype:
        return weight.detach()
    c = weight.__dict__.get("_sb200_half") if hasattr(weight, "__dict__") else None
    if c is not None and c[0].dtype == dtype and c[1] == weight._version:
        return c[0]
    return weight.detach().to(dtype)


# --------------------------------------------------------------------------------------------- sparse head
def head_forward(hidden, weight, bias, attention_mask, use_l0=False, want_aux=True, out=None, peer_ptrs=None):
    """Fused MLM-decoder GEMM + mask + max-pool + log1p(relu) (sparse_encoders.py:108-114).

    hidden [B,L,H] bf16, weight [V,H] bf16, bias [V] fp32 or None, attention_mask [B,L] (int64/int32/uint8/bool).
    out: optional preallocated contiguous fp32 [B,V] result buffer (e.g. this rank's slot of a PeerSink);
    peer_ptrs: up to 7 device pointers of the same [B,V] slot inside other ranks' gathered buffers -- the epilogue
    stores the result there too (the fused all-gather of gather_rep).
    Returns (rep [B,V] fp32, xmax [B,V] fp32 | None, argmax [B,V] int32 | None).
    """
    _need_cuda(hidden, weight, bias, attention_mask)
    if hidden.dtype != weight.dtype or hidden.dtype not in (torch.bfloat16, torch.float16):
        raise TypeError("head_forward expects hidden and weight both bf16 or both fp16")
    half_flag = _lib.HEAD_FP16 if hidden.dtype == torch.float16 else 0
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
    if out is None:
        rep = torch.empty(B, V, dtype=torch.float32, device=dev)
    else:
        if out.dtype != torch.float32 or tuple(out.shape) != (B, V) or not out.is_contiguous() or out.device != dev:
            raise ValueError("head_forward: `out` must be a contiguous fp32 [B, V] tensor on the inputs' device")
        rep = out
    xmax = torch.empty(B, V, dtype=torch.float32, device=dev) if want_aux else None
    argmax = torch.empty(B, V, dtype=torch.int32, device=dev) if want_aux else None
    ws_bytes = lib.sb200_head_fwd_workspace_bytes(B, L)
    ws = _workspace(ws_bytes, dev)
    peers = list(peer_ptrs or [])
    c_peers = (_lib.ctypes.c_void_p * max(1, len(peers)))(*peers) if peers else None
    with torch.cuda.device(dev), _timed("head_fwd"):
        code = lib.sb200_head_fwd(_ptr(hidden), _ptr(weight), _ptr(bias), _ptr(mask), mask.element_size(), B, L, H, V,
                                  (_lib.HEAD_L0 if use_l0 else 0) | half_flag, _ptr(rep), _ptr(xmax), _ptr(argmax), c_peers,
                                  len(peers), _ptr(ws), ws.numel(), _stream())
    _lib.check(code, "sb200_head_fwd")
    return rep, xmax, argmax


def head_backward(d_rep, xmax, argmax, hidden, weight, use_l0=False, want_bias_grad=True):
    """Sparse head backward -> (d_hidden [B,L,H] fp32, dW [V,H] fp32, dbias [V] fp32 | None). hidden / weight are the
    half-precision operands of the forward call (both bf16 or both fp16)."""
    _need_cuda(d_rep, xmax, argmax, hidden, weight)
    if hidden.dtype != weight.dtype or hidden.dtype not in (torch.bfloat16, torch.float16):
        raise TypeError("head_backward expects hidden and weight both bf16 or both fp16")
    half_flag = _lib.HEAD_FP16 if hidden.dtype == torch.float16 else 0
    B, L, H = hidden.shape
    V = weight.shape[0]
    lib = _lib.load()
    dev = hidden.device
    d_rep = d_rep.float().contiguous()
    d_hidden = torch.empty(B, L, H, dtype=torch.float32, device=dev)
    dW = torch.empty(V, H, dtype=torch.float32, device=dev)
    dbias = torch.empty(V, dtype=torch.float32, device=dev) if want_bias_grad else None
    ws = _workspace(lib.sb200_head_bwd_workspace_bytes(B, L, H, V), dev)
    with torch.cuda.device(dev), _timed("head_bwd"):
        code = lib.sb200_head_bwd(_ptr(d_rep), _ptr(xmax), _ptr(argmax), _ptr(hidden), _ptr(weight), B, L, H, V,
                                  (_lib.HEAD_L0 if use_l0 else 0) | half_flag, _ptr(d_hidden), _ptr(dW), _ptr(dbias), _ptr(ws),
                                  ws.numel(), _stream())
    _lib.check(code, "sb200_head_bwd")
    return d_hidden, dW, dbias


# --------------------------------------------------------------------------------------------- sparse head, packed input
def head_packed_supported(hidden_size, max_len):
    return bool(_lib.load().sb200_head_packed_supported(int(hidden_size), int(max_len)))


def _half_flag(hidden, weight, what):
    if hidden.dtype != weight.dtype or hidden.dtype not in (torch.bfloat16, torch.float16):
        raise TypeError(f"{what} expects hidden and weight both bf16 or both fp16")
    return _lib.HEAD_FP16 if hidden.dtype == torch.float16 else 0


def head_forward_packed(hidden, cu_seqlens, max_len, weight, bias, use_l0=False, want_aux=True, out=None, peer_ptrs=None):
    """The fused head on packed hidden states: hidden [T,H] (real tokens only), sequence b = rows
    [cu_seqlens[b], cu_seqlens[b+1]) (int32 [B+1] on the device). -> (rep [B,V], xmax | None, argmax | None); argmax is
    the token's rank inside its sequence."""
    _need_cuda(hidden, weight, bias, cu_seqlens)
    half = _half_flag(hidden, weight, "head_forward_packed")
    T, H = hidden.shape
    V = weight.shape[0]
    B = cu_seqlens.numel() - 1
    cu = cu_seqlens.to(torch.int32).contiguous()
    hidden, weight = hidden.contiguous(), weight.contiguous()
    if bias is not None:
        bias = bias.detach().float().contiguous()
    lib = _lib.load()
    dev = hidden.device
    if out is None:
        rep = torch.empty(B, V, dtype=torch.float32, device=dev)
    else:
        if out.dtype != torch.float32 or tuple(out.shape) != (B, V) or not out.is_contiguous():
            raise ValueError("head_forward_packed: `out` must be a contiguous fp32 [B, V] tensor")
        rep = out
    xmax = torch.empty(B, V, dtype=torch.float32, device=dev) if want_aux else None
    argmax = torch.empty(B, V, dtype=torch.int32, device=dev) if want_aux else None
    ws = _workspace(lib.sb200_head_fwd_workspace_bytes(B, int(max_len)), dev)
    peers = list(peer_ptrs or [])
    c_peers = (_lib.ctypes.c_void_p * max(1, len(peers)))(*peers) if peers else None
    with torch.cuda.device(dev), _timed("head_fwd"):
        code = lib.sb200_head_fwd_packed(_ptr(hidden), _ptr(weight), _ptr(bias), _ptr(cu), T, B, int(max_len), H, V,
                                         (_lib.HEAD_L0 if use_l0 else 0) | half, _ptr(rep), _ptr(xmax), _ptr(argmax),
                                         c_peers, len(peers), _ptr(ws), ws.numel(), _stream())
    _lib.check(code, "sb200_head_fwd_packed")
    return rep, xmax, argmax


def head_backward_packed(d_rep, xmax, argmax, hidden, cu_seqlens, max_len, weight, use_l0=False, want_bias_grad=True):
    """-> (d_hidden [T,H] fp32 packed, dW [V,H] fp32, dbias [V] fp32 | None)"""
    _need_cuda(d_rep, xmax, argmax, hidden, weight, cu_seqlens)
    half = _half_flag(hidden, weight, "head_backward_packed")
    T, H = hidden.shape
    V = weight.shape[0]
    B = cu_seqlens.numel() - 1
    lib = _lib.load()
    dev = hidden.device
    d_rep = d_rep.float().contiguous()
    d_hidden = torch.empty(T, H, dtype=torch.float32, device=dev)
    dW = torch.empty(V, H, dtype=torch.float32, device=dev)
    dbias = torch.empty(V, dtype=torch.float32, device=dev) if want_bias_grad else None
    ws = _workspace(lib.sb200_head_bwd_workspace_bytes(B, int(max_len), H, V), dev)
    with torch.cuda.device(dev), _timed("head_bwd"):
        code = lib.sb200_head_bwd_packed(_ptr(d_rep), _ptr(xmax), _ptr(argmax), _ptr(hidden), _ptr(weight),
                                         _ptr(cu_seqlens), T, B, int(max_len), H, V, (_lib.HEAD_L0 if use_l0 else 0) | half,
                                         _ptr(d_hidden), _ptr(dW), _ptr(dbias), _ptr(ws), ws.numel(), _stream())
    _lib.check(code, "sb200_head_bwd_packed")
    return d_hidden, dW, dbias


class SparseHeadPackedFunction(torch.autograd.Function):
    """SparseHeadFunction on the packed [T,H] activations of the padding-free body (no padded copy in either direction)."""

    @staticmethod
    def forward(ctx, hidden, cu_seqlens, max_len, weight, bias, use_l0, sink=None):
        half = torch.bfloat16
        if hidden.dtype == torch.float16 or (torch.is_autocast_enabled("cuda")
                                              and torch.get_autocast_dtype("cuda") == torch.float16):
            half = torch.float16
        h16 = hidden.detach().to(half).contiguous()
        w16 = half_weight(weight, half).contiguous()
        cu = cu_seqlens.to(torch.int32).contiguous()
        needs_grad = any(t is not None and t.requires_grad for t in (hidden, weight, bias))
        out = peers = None
        B = cu.numel() - 1
        if sink is not None and (sink.rows, sink.width) == (B, w16.shape[0]) and sink.world <= 8:
            out, peers = sink.slot(), sink.remote_slots()
        rep, xmax, argmax = head_forward_packed(h16, cu, max_len, w16, bias, use_l0, want_aux=needs_grad, out=out,
                                                peer_ptrs=peers)
        if needs_grad:
            ctx.save_for_backward(h16, w16, xmax, argmax, cu)
        ctx.use_l0, ctx.max_len = bool(use_l0), int(max_len)
        ctx.has_bias = bias is not None
        ctx.in_dtypes = (hidden.dtype, weight.dtype, None if bias is None else bias.dtype)
        return rep

    @staticmethod
    def backward(ctx, d_rep):
        h16, w16, xmax, argmax, cu = ctx.saved_tensors
        need_h, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[3], ctx.needs_input_grad[4]
        d_hidden, dW, dbias = head_backward_packed(d_rep, xmax, argmax, h16, cu, ctx.max_len, w16, ctx.use_l0,
                                                   want_bias_grad=ctx.has_bias)
        hd, wd, bd = ctx.in_dtypes
        return (d_hidden.to(hd) if need_h else None, None, None, dW.to(wd) if need_w else None,
                dbias.to(bd) if (need_b and ctx.has_bias) else None, None, None)


def sparse_head_packed(hidden, cu_seqlens, max_len, weight, bias, use_l0=False, sink=None):
    return SparseHeadPackedFunction.apply(hidden, cu_seqlens, max_len, weight, bias, use_l0, sink)


def prune_rows_(rep, ratio):
    """In-place row pruning (sparse_encoders.py:115-119)."""
    _need_cuda(rep)
    if rep.dtype != torch.float32 or not rep.is_contiguous():
        raise TypeError("prune_rows_ expects a contiguous fp32 matrix")
    with torch.cuda.device(rep.device):
        code = _lib.load().sb200_prune_rows(_ptr(rep), rep.shape[0], rep.shape[1], float(ratio), _stream())
    _lib.check(code, "sb200_prune_rows")
    return rep


class SparseHeadFunction(torch.autograd.Function):
    """rep = head(hidden, weight, bias, mask); gradients flow to hidden, weight and bias.
    Operand precision follows the reference's autocast: fp16 operands when the hidden states arrive in fp16 or fp16
    autocast is active (configs with `fp16: true` -- the decoder Linear then multiplies fp16 x fp16 upstream too), bf16
    otherwise (bf16 autocast, and fp32 callers: the tensor-core kernel has no fp32-operand mode; accumulation is fp32)."""

    @staticmethod
    def forward(ctx, hidden, weight, bias, attention_mask, use_l0, sink=None):
        half = torch.bfloat16
        if hidden.dtype == torch.float16 or (hidden.is_cuda and torch.is_autocast_enabled("cuda")
                                              and torch.get_autocast_dtype("cuda") == torch.float16):
            half = torch.float16
        h16 = hidden.detach().to(half).contiguous()
        w16 = half_weight(weight, half).contiguous()
        needs_grad = any(t is not None and t.requires_grad for t in (hidden, weight, bias))
        out = peers = None
        if sink is not None and (sink.rows, sink.width) == (h16.shape[0], w16.shape[0]) and sink.world <= 8:
            out, peers = sink.slot(), sink.remote_slots()    # fused all-gather: results land in every rank's buffer
        rep, xmax, argmax = head_forward(h16, w16, bias, attention_mask, use_l0, want_aux=needs_grad, out=out,
                                         peer_ptrs=peers)
        if needs_grad:
            ctx.save_for_backward(h16, w16, xmax, argmax)
        ctx.use_l0 = bool(use_l0)
        ctx.has_bias = bias is not None
        ctx.in_dtypes = (hidden.dtype, weight.dtype, None if bias is None else bias.dtype)
        return rep

    @staticmethod
    def backward(ctx, d_rep):
        h16, w16, xmax, argmax = ctx.saved_tensors
        need_h, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        d_hidden, dW, dbias = head_backward(d_rep, xmax, argmax, h16, w16, ctx.use_l0, want_bias_grad=ctx.has_bias)
        hd, wd, bd = ctx.in_dtypes
        return (d_hidden.to(hd) if need_h else None, dW.to(wd) if need_w else None,
                dbias.to(bd) if (need_b and ctx.has_bias) else None, None, None, None)


def sparse_head(hidden, weight, bias, attention_mask, use_l0=False, sink=None):
    """sink: an optional scripts.peer.PeerSink of shape [B, V] -- the head then writes its output rows straight into
    every rank's gathered buffer (gather_rep fused into the GEMM epilogue); follow with peer_gather(rep, sink)."""
    return SparseHeadFunction.apply(hidden, weight, bias, attention_mask, use_l0, sink)


# --------------------------------------------------------------------------------------------- inf-free query
def idf_query_forward(input_ids, idf_vector, special_ids, check_ids=None):
    """q[b, v] = relu(idf[v]) for the non-special token ids of row b (sparse_encoders.py:121-127), bit-exact.
    Token ids outside [0, V) make the reference raise an index error; here they are counted on the device and, when
    `check_ids` is true (default: whenever gradients are off, i.e. encode / eval, and no stream capture is running),
    the count is read back and an IndexError is raised. The training step skips the read-back (it would be a host
    synchronisation per step) -- such ids contribute nothing there."""
    _need_cuda(input_ids, idf_vector, special_ids)
    ids = input_ids
    if ids.dtype not in (torch.int64, torch.int32):
        ids = ids.to(torch.int64)
    ids = ids.contiguous()
    Nq, Lq = ids.shape
    idf = idf_vector.detach().float().contiguous()
    V = idf.numel()
    q = torch.empty(Nq, V, dtype=torch.float32, device=ids.device)
    if check_ids is None:
        check_ids = not torch.is_grad_enabled() and not torch.cuda.is_current_stream_capturing()
    bad = torch.empty(1, dtype=torch.int32, device=ids.device) if check_ids else None
    with torch.cuda.device(ids.device):
        code = _lib.load().sb200_idf_query(_ptr(ids), ids.element_size(), _ptr(idf), _ptr(special_ids),
                                           int(special_ids.numel()), Nq, Lq, V, _ptr(q), _ptr(bad), _stream())
    _lib.check(code, "sb200_idf_query")
    if bad is not None and int(bad) != 0:
        raise IndexError(f"idf_query: {int(bad)} token id(s) outside [0, {V})")
    return q


class IdfQueryFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input_ids, idf_vector, special_ids):
        q = idf_query_forward(input_ids, idf_vector, special_ids)
        if idf_vector.requires_grad:
            ctx.save_for_backward(q)
        ctx.idf_dtype = idf_vector.dtype
        # a row holds at most one entry per token: lets the score kernels skip their dense-query fallback launches
        q._sb200_nnz_bound = int(input_ids.shape[1])
        return q

    @staticmethod
    def backward(ctx, d_q):
        (q,) = ctx.saved_tensors
        d_q = d_q.float().contiguous()
        d_idf = torch.empty(q.shape[1], dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            code = _lib.load().sb200_idf_query_bwd(_ptr(d_q), _ptr(q), q.shape[0], q.shape[1], _ptr(d_idf), _stream())
        _lib.check(code, "sb200_idf_query_bwd")
        return None, d_idf.to(ctx.idf_dtype), None


def idf_query(input_ids, idf_vector, special_ids):
    return IdfQueryFunction.apply(input_ids, idf_vector, special_ids)


# --------------------------------------------------------------------------------------------- FLOPS regulariser
def flops_forward(rep, group_num=1, threshold=None, want_stats=False):
    """-> (value [] fp32, colsum [G,V], rowmask [N*G] | None, stats [4] | None). The row mask exists only for the
    L0-thresholded variant (or with stats); the plain regulariser is one launch over rep."""
    _need_cuda(rep)
    rep = rep.float().contiguous()
    rows, V = rep.shape
    if rows % group_num != 0:
        raise ValueError(f"{rows} rows are not divisible by group_num={group_num}")
    N = rows // group_num
    dev = rep.device
    colsum = torch.empty(group_num, V, dtype=torch.float32, device=dev)
    need_rows = threshold is not None or want_stats
    rowmask = torch.empty(rows, dtype=torch.float32, device=dev) if need_rows else None
    value = torch.empty((), dtype=torch.float32, device=dev)
    stats = torch.empty(4, dtype=torch.float32, device=dev) if want_stats else None
    lib = _lib.load()
    ws = _workspace(lib.sb200_flops_workspace_bytes(N, group_num, V), dev)
    with torch.cuda.device(dev):
        code = lib.sb200_flops_fwd(_ptr(rep), N, group_num, V, -1.0 if threshold is None else float(threshold),
                                   _ptr(colsum), _ptr(rowmask), _ptr(value), _ptr(stats), _ptr(ws), ws.numel(), _stream())
    _lib.check(code, "sb200_flops_fwd")
    return value, colsum, rowmask, stats


class FlopsFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rep, group_num, threshold):
        rep32 = rep.detach().float().contiguous()
        value, colsum, rowmask, _ = flops_forward(rep32, group_num, threshold)
        ctx.save_for_backward(rep32, colsum, rowmask)
        ctx.group_num = group_num
        ctx.rep_dtype = rep.dtype
        ctx.grad_rows = getattr(rep, "_sb200_grad_rows", None)  # set by gather_rep: only these rows carry grad
        return value

    @staticmethod
    def backward(ctx, g):
        rep, colsum, rowmask = ctx.saved_tensors
        rows, V = rep.shape
        G = ctx.group_num
        g = g.detach().float().reshape(1).contiguous()
        lo, hi = ctx.grad_rows if ctx.grad_rows is not None else (0, rows)
        d_rep = torch.empty_like(rep) if (lo, hi) == (0, rows) else torch.zeros_like(rep)
        with torch.cuda.device(rep.device):
            code = _lib.load().sb200_flops_bwd(_ptr(rep), _ptr(colsum), _ptr(rowmask), _ptr(g), rows // G, G, V, lo, hi,
                                               0, _ptr(d_rep), _stream())
        _lib.check(code, "sb200_flops_bwd")
        return d_rep.to(ctx.rep_dtype), None, None


def flops_value(rep, group_num=1, threshold=None):
    """trainer.py:61-73"""
    shape_v = rep.shape[-1]
    flat = rep if rep.dim() == 2 else rep.reshape(-1, shape_v)
    return FlopsFunction.apply(flat, int(group_num), threshold)


# --------------------------------------------------------------------------------------------- scores + losses
def _check_qd(q, d, in_batch):
    _need_cuda(q, d)
    q = q.float().contiguous()
    d = d.float().contiguous()
    Nq, V = q.shape
    Nd = d.shape[0]
    if d.shape[1] != V:
        raise ValueError("q and d must share the vocabulary dimension")
    if not in_batch and Nd % Nq != 0:
        raise ValueError("number of docs must be a multiple of the number of queries")
    return q, d, Nq, Nd, V


def scores_forward(q, d, in_batch, return_workspace=False, q_nnz_bound=None):
    """q . d^T. q_nnz_bound: promised maximum of non-zeros per query row (default: what ops.idf_query recorded on q)."""
    if q_nnz_bound is None:
        q_nnz_bound = int(getattr(q, "_sb200_nnz_bound", 0) or 0)
    q, d, Nq, Nd, V = _check_qd(q, d, in_batch)
    C = Nd if in_batch else Nd // Nq
    S = torch.empty(Nq, C, dtype=torch.float32, device=q.device)
    lib = _lib.load()
    ws = _workspace(lib.sb200_scores_workspace_bytes(Nq, Nd, V, 1 if in_batch else 0), q.device)
    with torch.cuda.device(q.device):
        code = lib.sb200_scores_fwd(_ptr(q), _ptr(d), Nq, Nd, V, 1 if in_batch else 0, int(q_nnz_bound), _ptr(S), _ptr(ws),
                                    ws.numel(), _stream())
    _lib.check(code, "sb200_scores_fwd")
    return (S, ws) if return_workspace else S


def _scores_backward(dS, gscale, q, d, in_batch, q_rows, d_rows, need_q, need_d, ws):
    """(d_q | None, d_d | None) = gscale * (dS d, dS^T q) restricted to the row ranges that carry gradient."""
    Nq, V = q.shape
    Nd = d.shape[0]
    q_lo, q_hi = q_rows if q_rows is not None else (0, Nq)
    d_lo, d_hi = d_rows if d_rows is not None else (0, Nd)
    d_q = (torch.empty_like(q) if (q_lo, q_hi) == (0, Nq) else torch.zeros_like(q)) if need_q else None
    d_d = (torch.empty_like(d) if (d_lo, d_hi) == (0, Nd) else torch.zeros_like(d)) if need_d else None
    with torch.cuda.device(q.device):
        code = _lib.load().sb200_scores_bwd(_ptr(dS), _ptr(gscale), _ptr(q), _ptr(d), Nq, Nd, V, 1 if in_batch else 0,
                                            q_lo, q_hi, d_lo, d_hi, 0, _ptr(d_q), _ptr(d_d), _ptr(ws),
                                            0 if ws is None else ws.numel(), _stream())
    _lib.check(code, "sb200_scores_bwd")
    return d_q, d_d


class ScoresFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, d, in_batch):
        q32 = q.detach().float().contiguous()
        d32 = d.detach().float().contiguous()
        S, ws = scores_forward(q32, d32, in_batch, return_workspace=True,
                               q_nnz_bound=int(getattr(q, "_sb200_nnz_bound", 0) or 0))
        ctx.save_for_backward(q32, d32)
        ctx.ws = ws if in_batch else None  # thresholded query lists, reused by the backward kernels
        ctx.q_rows = getattr(q, "_sb200_grad_rows", None)  # set by gather_rep: only these rows carry grad
        ctx.d_rows = getattr(d, "_sb200_grad_rows", None)
        ctx.in_batch = bool(in_batch)
        ctx.dtypes = (q.dtype, d.dtype)
        return S

    @staticmethod
    def backward(ctx, dS):
        q, d = ctx.saved_tensors
        need_q, need_d = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_q, d_d = _scores_backward(dS.float().contiguous(), None, q, d, ctx.in_batch, ctx.q_rows, ctx.d_rows, need_q,
                                    need_d, ctx.ws)
        return (d_q.to(ctx.dtypes[0]) if need_q else None, d_d.to(ctx.dtypes[1]) if need_d else None, None)


def scores(q, d, in_batch):
    """q . d^T (all pairs) or each query against its own docs (loss.py:30-37)."""
    return ScoresFunction.apply(q, d, in_batch)


_LOSS_MODES = {"infonce": _lib.LOSS_INFONCE, "kldiv": _lib.LOSS_KLDIV, "marginmse": _lib.LOSS_MARGINMSE}


class RankLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, S, teacher, mode, G, in_batch, temperature):
        _need_cuda(S, teacher)
        S32 = S.detach().float().contiguous()
        Nq, C = S32.shape
        t32 = None
        if teacher is not None:
            t32 = teacher.detach().float().contiguous()
            if tuple(t32.shape) != (Nq, C):
                raise ValueError(f"teacher scores {tuple(t32.shape)} do not match student scores {(Nq, C)}")
        loss = torch.empty((), dtype=torch.float32, device=S.device)
        dS = torch.empty_like(S32) if S.requires_grad else None
        lib = _lib.load()
        ws = _workspace(lib.sb200_rank_loss_workspace_bytes(Nq), S.device)
        with torch.cuda.device(S.device):
            code = lib.sb200_rank_loss(_LOSS_MODES[mode], _ptr(S32), _ptr(t32), Nq, C, int(G), 1 if in_batch else 0,
                                       float(temperature), _ptr(loss), _ptr(dS), _ptr(ws), ws.numel(), _stream())
        _lib.check(code, "sb200_rank_loss")
        if dS is not None:
            ctx.save_for_backward(dS)
        ctx.s_dtype = S.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        (dS,) = ctx.saved_tensors
        return (dS * g).to(ctx.s_dtype), None, None, None, None, None


def rank_loss(S, teacher, mode, G, in_batch, temperature=1.0):
    return RankLossFunction.apply(S, teacher, mode, G, in_batch, temperature)


def score_loss_forward(q, d, teacher, mode, G, in_batch, temperature=1.0, q_nnz_bound=0, want_grad=True):
    """Scores + ranking loss in one call (sb200_score_loss_fwd): -> (loss [], S [Nq,C], dS [Nq,C] | None, workspace)."""
    q, d, Nq, Nd, V = _check_qd(q, d, in_batch)
    if Nd != Nq * int(G):
        raise ValueError(f"{Nd} docs != {Nq} queries x G={G}")
    C = Nd if in_batch else Nd // Nq
    t32 = None
    if teacher is not None:
        _need_cuda(teacher)
        t32 = teacher.detach().float().contiguous()
        if tuple(t32.shape) != (Nq, C):
            raise ValueError(f"teacher scores {tuple(t32.shape)} do not match student scores {(Nq, C)}")
    S = torch.empty(Nq, C, dtype=torch.float32, device=q.device)
    dS = torch.empty(Nq, C, dtype=torch.float32, device=q.device) if want_grad else None
    loss = torch.empty((), dtype=torch.float32, device=q.device)
    lib = _lib.load()
    ws = _workspace(lib.sb200_scores_workspace_bytes(Nq, Nd, V, 1 if in_batch else 0), q.device)
    with torch.cuda.device(q.device):
        code = lib.sb200_score_loss_fwd(_LOSS_MODES[mode], _ptr(q), _ptr(d), _ptr(t32), Nq, Nd, V, int(G),
                                        1 if in_batch else 0, float(temperature), int(q_nnz_bound), _ptr(S), _ptr(loss),
                                        _ptr(dS), _ptr(ws), ws.numel(), _stream())
    _lib.check(code, "sb200_score_loss_fwd")
    return loss, S, dS, ws


class ScoreLossFunction(torch.autograd.Function):
    """loss = ranking_loss(q . d^T [, teacher]) with the score matrix, the loss and d loss / d S produced by one fused
    forward call; the backward multiplies by the upstream scalar inside the score-backward kernels."""

    @staticmethod
    def forward(ctx, q, d, teacher, mode, G, in_batch, temperature):
        q32 = q.detach().float().contiguous()
        d32 = d.detach().float().contiguous()
        want = q.requires_grad or d.requires_grad
        loss, S, dS, ws = score_loss_forward(q32, d32, teacher, mode, G, in_batch, temperature,
                                             q_nnz_bound=int(getattr(q, "_sb200_nnz_bound", 0) or 0), want_grad=want)
        if want:
            ctx.save_for_backward(q32, d32, dS)
        ctx.ws = ws if in_batch else None
        ctx.q_rows = getattr(q, "_sb200_grad_rows", None)
        ctx.d_rows = getattr(d, "_sb200_grad_rows", None)
        ctx.in_batch = bool(in_batch)
        ctx.dtypes = (q.dtype, d.dtype)
        ctx.mark_non_differentiable(S)
        return loss, S

    @staticmethod
    def backward(ctx, g, _g_scores):
        q, d, dS = ctx.saved_tensors
        need_q, need_d = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gscale = g.detach().float().reshape(1).contiguous()
        d_q, d_d = _scores_backward(dS, gscale, q, d, ctx.in_batch, ctx.q_rows, ctx.d_rows, need_q, need_d, ctx.ws)
        return (d_q.to(ctx.dtypes[0]) if need_q else None, d_d.to(ctx.dtypes[1]) if need_d else None, None, None, None,
                None, None)


def score_loss(q, d, teacher, mode, G, in_batch, temperature=1.0):
    """-> scalar loss (loss.py:25-43, 57-77, 86-107 on dense q_rep / d_rep)."""
    return ScoreLossFunction.apply(q, d, teacher, mode, int(G), bool(in_batch), float(temperature))[0]


# --------------------------------------------------------------------------------------------- encode output path
def compact_rows(rep, first_col=1, df_count=None, capacity=None):
    """Dense [B,V] -> CSR (row_ptr [B+1] i32, cols i32, vals f32) of the non-zero entries of columns >= first_col.
    df_count (int64 [V], optional) is incremented where rep > 0 (sparse_encoders.py:178-179)."""
    _need_cuda(rep, df_count)
    rep = rep.float().contiguous()
    B, V = rep.shape
    dev = rep.device
    cap = B * V if capacity is None else int(capacity)
    row_ptr = torch.empty(B + 1, dtype=torch.int32, device=dev)
    cols = torch.empty(cap, dtype=torch.int32, device=dev)
    vals = torch.empty(cap, dtype=torch.float32, device=dev)
    lib = _lib.load()
    ws = _workspace(lib.sb200_compact_workspace_bytes(B, V), dev)
    with torch.cuda.device(dev):
        code = lib.sb200_compact_rows(_ptr(rep), B, V, int(first_col), _ptr(row_ptr), _ptr(cols), _ptr(vals), cap,
                                      _ptr(df_count), _ptr(ws), ws.numel(), _stream())
    _lib.check(code, "sb200_compact_rows")
    return row_ptr, cols, vals


def minmax_accumulate(S, acc=None, scale=1.0):
    """acc (+)= scale * (S - rowmin) / (rowmax - rowmin + 1e-6)   (bi_encoder_wrapper.py:133-138)"""
    _need_cuda(S, acc)
    S = S.float().contiguous()
    out = torch.empty_like(S) if acc is None else acc
    with torch.cuda.device(S.device):
        code = _lib.load().sb200_minmax_accumulate(_ptr(S), S.shape[0], S.shape[1], float(scale), 0 if acc is None else 1,
                                                   _ptr(out), _stream())
    _lib.check(code, "sb200_minmax_accumulate")
    return out


# --------------------------------------------------------------------------------------------- encoder body: LayerNorm
def layer_norm_supported(hidden_size):
    return bool(_lib.load().sb200_layer_norm_supported(int(hidden_size)))


class LayerNormFunction(torch.autograd.Function):
    """y = LayerNorm(x) over the last dimension; x bf16 or fp32 (kept), gamma/beta fp32, statistics in fp32."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        _need_cuda(x, gamma, beta)
        H = x.shape[-1]
        xc = x.detach()
        if xc.dtype not in (torch.bfloat16, torch.float32):
            xc = xc.float()
        xc = xc.contiguous()
        R = xc.numel() // H
        g32 = gamma.detach().float().contiguous()
        b32 = beta.detach().float().contiguous()
        y = torch.empty_like(xc)
        mean = torch.empty(R, dtype=torch.float32, device=xc.device)
        rstd = torch.empty(R, dtype=torch.float32, device=xc.device)
        with torch.cuda.device(xc.device):
            code = _lib.load().sb200_layer_norm_fwd(_ptr(xc), xc.element_size(), _ptr(g32), _ptr(b32), R, H, float(eps),
                                                    _ptr(y), _ptr(mean), _ptr(rstd), _stream())
        _lib.check(code, "sb200_layer_norm_fwd")
        ctx.save_for_backward(xc, g32, mean, rstd)
        ctx.in_dtypes = (x.dtype, gamma.dtype, beta.dtype)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        xc, g32, mean, rstd = ctx.saved_tensors
        H = xc.shape[-1]
        R = xc.numel() // H
        dyc = dy.to(xc.dtype).contiguous()
        dx = torch.empty_like(xc)
        dgamma = torch.empty(H, dtype=torch.float32, device=xc.device)
        dbeta = torch.empty(H, dtype=torch.float32, device=xc.device)
        lib = _lib.load()
        ws = _workspace(lib.sb200_layer_norm_bwd_workspace_bytes(R, H), xc.device)
        with torch.cuda.device(xc.device):
            code = lib.sb200_layer_norm_bwd(_ptr(xc), _ptr(dyc), xc.element_size(), _ptr(g32), _ptr(mean), _ptr(rstd), R, H,
                                            _ptr(dx), _ptr(dgamma), _ptr(dbeta), _ptr(ws), ws.numel(), _stream())
        _lib.check(code, "sb200_layer_norm_bwd")
        xd, gd, bd = ctx.in_dtypes
        return dx.view(xc.shape).to(xd), dgamma.to(gd), dbeta.to(bd), None


def layer_norm(x, gamma, beta, eps=1e-12):
    return LayerNormFunction.apply(x, gamma, beta, eps)


def add_layer_norm_forward(y, resid, gamma, beta, eps, seed=None, p=0.0, want_f32=True, want_bf16=True):
    """Raw forward of the fused block tail LayerNorm(dropout(y) + resid): returns (out_f32, out_bf16, mean, rstd).
    y bf16 [..., H]; resid fp32 or None; seed: int64 device tensor (1 element) or None."""
    _need_cuda(y, resid, gamma, beta, seed)
    H = y.shape[-1]
    R = y.numel() // H
    dev = y.device
    out32 = torch.empty(y.shape, dtype=torch.float32, device=dev) if want_f32 else None
    out16 = torch.empty(y.shape, dtype=torch.bfloat16, device=dev) if want_bf16 else None
    mean = torch.empty(R, dtype=torch.float32, device=dev)
    rstd = torch.empty(R, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        code = _lib.load().sb200_add_layer_norm_fwd(_ptr(y), _ptr(resid), _ptr(gamma), _ptr(beta), R, H, float(eps),
                                                    _ptr(seed), float(p), _ptr(out32), _ptr(out16), _ptr(mean),
                                                    _ptr(rstd), _stream())
    _lib.check(code, "sb200_add_layer_norm_fwd")
    return out32, out16, mean, rstd


class AddLayerNormFunction(torch.autograd.Function):
    """(out_f32, out_bf16) = LayerNorm(dropout(y) + resid): the tail of a BERT attention / feed-forward block
    (transformers BertSelfOutput / BertOutput) under bf16 autocast, with the next GEMM's bf16 operand produced by
    the same kernel. The dropout mask is regenerated from `seed` in the backward pass."""

    @staticmethod
    def forward(ctx, y, resid, gamma, beta, eps, seed, p, want_f32):
        ctx.set_materialize_grads(False)
        yc = y.detach()
        if yc.dtype != torch.bfloat16:
            raise TypeError("add_layer_norm: the branch output must be bf16 (run the body under bf16 autocast)")
        yc = yc.contiguous()
        rc = None if resid is None else resid.detach().float().contiguous()
        g32 = gamma.detach().float().contiguous()
        b32 = beta.detach().float().contiguous()
        out32, out16, mean, rstd = add_layer_norm_forward(yc, rc, g32, b32, eps, seed, p, want_f32=want_f32)
        ctx.save_for_backward(yc, rc, g32, mean, rstd, seed)
        ctx.p = float(p)
        ctx.dtypes = (gamma.dtype, beta.dtype, None if resid is None else resid.dtype)
        return out32, out16

    @staticmethod
    def backward(ctx, g32_out, g16_out):
        yc, rc, gam, mean, rstd, seed = ctx.saved_tensors
        if g32_out is None and g16_out is None:
            return (None,) * 8
        H = yc.shape[-1]
        R = yc.numel() // H
        dev = yc.device
        if g32_out is not None:
            g32_out = g32_out.float().contiguous()
        if g16_out is not None:
            g16_out = g16_out.to(torch.bfloat16).contiguous()
        dy = torch.empty_like(yc)
        dres = torch.empty(yc.shape, dtype=torch.float32, device=dev) if rc is not None else None
        dgamma = torch.empty(H, dtype=torch.float32, device=dev)
        dbeta = torch.empty(H, dtype=torch.float32, device=dev)
        lib = _lib.load()
        ws = _workspace(lib.sb200_layer_norm_bwd_workspace_bytes(R, H), dev)
        with torch.cuda.device(dev):
            code = lib.sb200_add_layer_norm_bwd(_ptr(yc), _ptr(rc), _ptr(g32_out), _ptr(g16_out), _ptr(gam), _ptr(mean),
                                                _ptr(rstd), R, H, _ptr(seed), ctx.p, _ptr(dy), _ptr(dres), _ptr(dgamma),
                                                _ptr(dbeta), _ptr(ws), ws.numel(), _stream())
        _lib.check(code, "sb200_add_layer_norm_bwd")
        gd, bd, rd = ctx.dtypes
        return dy, (None if dres is None else dres.to(rd)), dgamma.to(gd), dbeta.to(bd), None, None, None, None


def add_layer_norm(y, resid, gamma, beta, eps=1e-12, p=0.0, training=False, want_f32=True):
    """LayerNorm(dropout(y, p) + resid) -> (fp32 output or None, bf16 output). Dropout only when training and p > 0:
    a fresh 64-bit seed is drawn on the device from torch's CUDA generator (CUDA-graph safe: the generator's
    offset is advanced by the graph on every replay)."""
    seed = None
    if training and p > 0.0:
        seed = torch.randint(-2 ** 62, 2 ** 62, (1,), dtype=torch.int64, device=y.device)
    return AddLayerNormFunction.apply(y, resid, gamma, beta, eps, seed, float(p) if seed is not None else 0.0, want_f32)


# --------------------------------------------------------------------------------------------- encoder body: embeddings
class EmbedSumFunction(torch.autograd.Function):
    """(W[ids] + T[typ]) + P[pos] on token rows, fp32 tables (transformers BertEmbeddings lookups); the backward
    scatters the gradient into the three dense table gradients with vector reductions (summation order is not
    deterministic, like the head backward)."""

    @staticmethod
    def forward(ctx, ids, pos, typ, W, P, T, padding_idx):
        _need_cuda(ids, pos, typ, W, P, T)
        if not (W.dtype == P.dtype == T.dtype == torch.float32):
            raise TypeError("embed_sum: fp32 embedding tables required")
        ids, pos, typ = (t.to(torch.int64).contiguous() for t in (ids, pos, typ))
        Wc, Pc, Tc = W.detach().contiguous(), P.detach().contiguous(), T.detach().contiguous()
        n, H = ids.numel(), W.shape[1]
        out = torch.empty(n, H, dtype=torch.float32, device=W.device)
        with torch.cuda.device(W.device):
            code = _lib.load().sb200_embed_sum_fwd(_ptr(ids), _ptr(pos), _ptr(typ), _ptr(Wc), _ptr(Pc), _ptr(Tc), n, H,
                                                   W.shape[0], P.shape[0], T.shape[0], _ptr(out), _stream())
        _lib.check(code, "sb200_embed_sum_fwd")
        ctx.save_for_backward(ids, pos, typ)
        ctx.shapes = (W.shape, P.shape, T.shape)
        ctx.padding_idx = -1 if padding_idx is None else int(padding_idx)
        return out

    @staticmethod
    def backward(ctx, g):
        ids, pos, typ = ctx.saved_tensors
        sw, sp, st = ctx.shapes
        g = g.float().contiguous()
        dW = torch.zeros(sw, dtype=torch.float32, device=g.device)
        dP = torch.zeros(sp, dtype=torch.float32, device=g.device)
        dT = torch.zeros(st, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            code = _lib.load().sb200_embed_sum_bwd(_ptr(ids), _ptr(pos), _ptr(typ), _ptr(g), ids.numel(), sw[1], sw[0],
                                                   sp[0], st[0], ctx.padding_idx, _ptr(dW), _ptr(dP), _ptr(dT), _stream())
        _lib.check(code, "sb200_embed_sum_bwd")
        return None, None, None, dW, dP, dT, None


def embed_sum(ids, pos, typ, word_table, pos_table, type_table, padding_idx=None):
    return EmbedSumFunction.apply(ids, pos, typ, word_table, pos_table, type_table, padding_idx)


# --------------------------------------------------------------------------------------------- encoder body: attention
def attn_supported(head_dim, max_len):
    return bool(_lib.load().sb200_attn_supported(int(head_dim), int(max_len)))


def _attn_views(qkv):
    """qkv bf16 [T, 3, h, d] (the fused projection output) -> (T, h, d, row stride in elements)."""
    if qkv.dtype != torch.bfloat16 or qkv.dim() != 4 or qkv.shape[1] != 3:
        raise TypeError("varlen_attention: qkv must be bf16 [T, 3, heads, head_dim]")
    T, _, h, d = qkv.shape
    return T, h, d, 3 * h * d


def attn_forward(qkv, cu_seqlens, max_len, scale, drop_p=0.0, seed=None, salt=0, live=None):
    """Raw forward: returns (out bf16 [T, h, d], lse f32 [h, T]). live: number of leading real sequences (the rest
    are filler: zero output rows, nothing computed); None = all."""
    _need_cuda(qkv, cu_seqlens, seed)
    T, h, d, stride = _attn_views(qkv)
    if cu_seqlens.dtype != torch.int32:
        raise TypeError("varlen_attention: cu_seqlens must be int32")
    qkv = qkv.contiguous()
    out = torch.empty(T, h, d, dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty(h, T, dtype=torch.float32, device=qkv.device)
    base, e = qkv.data_ptr(), 2 * h * d
    with torch.cuda.device(qkv.device):
        nseq = cu_seqlens.numel() - 1
        code = _lib.load().sb200_attn_fwd(base, base + e, base + 2 * e, stride, _ptr(cu_seqlens), nseq,
                                          nseq if live is None else int(live), int(max_len), T, h, d, float(scale), float(drop_p),
                                          _ptr(seed), int(salt), _ptr(out), _ptr(lse), _stream())
    _lib.check(code, "sb200_attn_fwd")
    return out, lse


def attn_backward(qkv, out, dout, lse, cu_seqlens, max_len, scale, drop_p=0.0, seed=None, salt=0, zero_fill=False,
                  live=None):
    """Raw backward: returns dqkv bf16 [T, 3, h, d]. Rows outside every sequence are only defined with zero_fill."""
    T, h, d, stride = _attn_views(qkv)
    dout = dout.contiguous()
    if dout.dtype != torch.bfloat16:
        dout = dout.to(torch.bfloat16)
    dqkv = (torch.zeros if zero_fill else torch.empty)(T, 3, h, d, dtype=torch.bfloat16, device=qkv.device)
    dsum = torch.empty(2, h, T, dtype=torch.float32, device=qkv.device)
    base, dbase, e = qkv.data_ptr(), dqkv.data_ptr(), 2 * h * d
    with torch.cuda.device(qkv.device):
        nseq = cu_seqlens.numel() - 1
        code = _lib.load().sb200_attn_bwd(base, base + e, base + 2 * e, stride, _ptr(out), _ptr(dout), _ptr(lse),
                                          _ptr(cu_seqlens), nseq, nseq if live is None else int(live), int(max_len), T, h, d,
                                          float(scale), float(drop_p), _ptr(seed), int(salt), dbase, dbase + e,
                                          dbase + 2 * e, stride, _ptr(dsum), _stream())
    _lib.check(code, "sb200_attn_bwd")
    return dqkv


def attn_dropout_mask(cu_seqlens, max_len, T, h, drop_p, seed, salt=0):
    """Test hook: bool [h, T, max_len], entry [head, t, j] = query t keeps key j of its own sequence."""
    _need_cuda(cu_seqlens, seed)
    mask = torch.zeros(h, T, int(max_len), dtype=torch.uint8, device=cu_seqlens.device)
    with torch.cuda.device(cu_seqlens.device):
        code = _lib.load().sb200_attn_dropout_mask(_ptr(cu_seqlens), cu_seqlens.numel() - 1, int(max_len), T, h,
                                                   float(drop_p), _ptr(seed), int(salt), _ptr(mask), _stream())
    _lib.check(code, "sb200_attn_dropout_mask")
    return mask.bool()


class VarlenAttentionFunction(torch.autograd.Function):
    """softmax(Q K^T * scale) -> dropout -> V per packed sequence (transformers BertSelfAttention under bf16 autocast)."""

    @staticmethod
    def forward(ctx, qkv, cu_seqlens, max_len, scale, drop_p, seed, salt, covers_all_rows, live):
        qc = qkv.detach().contiguous()
        out, lse = attn_forward(qc, cu_seqlens, max_len, scale, drop_p, seed, salt, live)
        ctx.save_for_backward(qc, out, lse, cu_seqlens, seed)
        ctx.cfg = (int(max_len), float(scale), float(drop_p), int(salt), bool(covers_all_rows), live)
        return out

    @staticmethod
    def backward(ctx, dout):
        qc, out, lse, cu, seed = ctx.saved_tensors
        max_len, scale, drop_p, salt, covers, live = ctx.cfg
        dqkv = attn_backward(qc, out, dout, lse, cu, max_len, scale, drop_p, seed, salt, zero_fill=not covers, live=live)
        return dqkv, None, None, None, None, None, None, None, None


def varlen_attention(qkv, cu_seqlens, max_len, scale=None, drop_p=0.0, training=False, seed=None, salt=0,
                     covers_all_rows=False, live_sequences=None):
    """Self-attention over packed sequences. qkv bf16 [T, 3, h, d]; cu_seqlens int32 [nseq + 1] (device); returns bf16
    [T, h, d] (rows outside every sequence are undefined). Dropout only when training and drop_p > 0; `seed` is a
    1-element int64 device tensor (drawn from torch's CUDA generator when None: CUDA-graph safe), `salt` separates
    calls that share a seed (the layer index). covers_all_rows: cu_seqlens[-1] == T, so the gradient needs no
    zero fill. live_sequences: only the first that many sequences are real; the others (filler rows of a packed
    batch) get zero output / gradient rows without any attention being computed."""
    d = qkv.shape[-1]
    if scale is None:
        scale = 1.0 / (d ** 0.5)
    p = float(drop_p) if training else 0.0
    if p > 0.0 and seed is None:
        seed = torch.randint(-2 ** 62, 2 ** 62, (1,), dtype=torch.int64, device=qkv.device)
    if p == 0.0:
        seed = None
    return VarlenAttentionFunction.apply(qkv, cu_seqlens, max_len, scale, p, seed, salt, covers_all_rows, live_sequences)


# --------------------------------------------------------------------------------------------- encoder body: Linear
def colsum_supported(n):
    return bool(_lib.load().sb200_colsum_supported(int(n)))


def colsum(dy):
    """Column sums of a 2-D bf16/fp32 matrix in fp32 (the bias gradient of a Linear layer)."""
    _need_cuda(dy)
    if dy.dtype not in (torch.bfloat16, torch.float32):
        dy = dy.float()
    dy = dy.contiguous()
    R, N = dy.shape
    out = torch.empty(N, dtype=torch.float32, device=dy.device)
    lib = _lib.load()
    ws = _workspace(lib.sb200_colsum_workspace_bytes(R, N), dy.device)
    with torch.cuda.device(dy.device):
        code = lib.sb200_colsum(_ptr(dy), dy.element_size(), R, N, _ptr(out), _ptr(ws), ws.numel(), _stream())
    _lib.check(code, "sb200_colsum")
    return out


def _weight_grad(dy2, x2, w_dtype):
    """dy2^T @ x2 in the dtype of the weight the gradient belongs to."""
    if dy2.dtype != x2.dtype:
        dy2 = dy2.to(x2.dtype)
    if w_dtype == torch.float32 and x2.dtype in (torch.bfloat16, torch.float16):
        return torch.mm(dy2.t(), x2, out_dtype=torch.float32)
    return (dy2.t() @ x2).to(w_dtype)


class LinearFunction(torch.autograd.Function):
    """y = x W^T + b with the library GEMMs of torch (cuBLAS) and the fused column-sum kernel for the bias gradient.
    Operands arrive already in the compute dtype (see `linear`)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        # the fp32 master weight is the autograd input: its gradient leaves the weight-gradient GEMM in fp32
        # (bf16 operands, fp32 output) instead of being rounded to bf16 and cast back
        w = half_weight(weight, x.dtype)
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        ctx.w_dtype = weight.dtype
        # the bias keeps its own (fp32) dtype as an autograd input so that its gradient stays the fp32 column sum
        return torch.nn.functional.linear(x, w, None if bias is None else half_weight(bias, x.dtype))

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = (dy2 @ weight).view(x.shape)
        if ctx.needs_input_grad[1]:
            dw = _weight_grad(dy2, x.reshape(-1, x.shape[-1]), ctx.w_dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy2)
        return dx, dw, db


class FusedQKVFunction(torch.autograd.Function):
    """[q | k | v] = x [Wq; Wk; Wv]^T + [bq; bk; bv] as ONE GEMM (transformers BertSelfAttention.query / key / value on
    the same activation). The concatenation happens on the cached half-precision copies (half the bytes of the fp32
    concatenation + cast it replaces); the fp32 master parameters are the autograd inputs and receive slices of one
    fp32 weight-gradient GEMM and of one column-sum."""

    @staticmethod
    def forward(ctx, x, wq, wk, wv, bq, bk, bv):
        dt = x.dtype
        w = torch.cat([half_weight(wq, dt), half_weight(wk, dt), half_weight(wv, dt)], 0)
        b = torch.cat([half_weight(bq, dt), half_weight(bk, dt), half_weight(bv, dt)], 0)
        ctx.save_for_backward(x, w)
        ctx.sizes = (wq.shape[0], wk.shape[0], wv.shape[0])
        ctx.w_dtype = wq.dtype
        return torch.nn.functional.linear(x, w, b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if dy2.dtype != x.dtype:
            dy2 = dy2.to(x.dtype)
        dx = (dy2 @ w).view(x.shape) if ctx.needs_input_grad[0] else None
        dw = _weight_grad(dy2, x.reshape(-1, x.shape[-1]), ctx.w_dtype).split(ctx.sizes, 0)
        db = colsum(dy2).split(ctx.sizes, 0)
        return (dx,) + tuple(dw) + tuple(db)


def fused_qkv(x, wq, wk, wv, bq, bk, bv):
    """The three attention projections of one activation as one GEMM; autocast semantics as `linear`."""
    if torch.is_autocast_enabled("cuda"):
        dt = torch.get_autocast_dtype("cuda")
        with torch.autocast("cuda", enabled=False):
            return FusedQKVFunction.apply(x.to(dt), wq, wk, wv, bq, bk, bv)
    return FusedQKVFunction.apply(x, wq, wk, wv, bq, bk, bv)


def gelu_forward(x):
    """Exact GELU of a bf16 tensor (numel % 8 == 0) on the sm_100a kernel."""
    _need_cuda(x)
    if x.dtype != torch.bfloat16 or x.numel() % 8 != 0 or x.numel() == 0:
        raise TypeError("gelu: bf16 input with a positive multiple of 8 elements required")
    x = x.contiguous()
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        code = _lib.load().sb200_gelu_fwd(_ptr(x), x.numel(), _ptr(y), _stream())
    _lib.check(code, "sb200_gelu_fwd")
    return y


def gelu_backward(x, dy, want_colsum=True):
    """(dx bf16 [R, N], colsum fp32 [N] or None) for y = gelu(x); colsum = column sums of dx (bias gradient)."""
    _need_cuda(x, dy)
    N = x.shape[-1]
    R = x.numel() // N
    x = x.contiguous()
    dy = dy.to(torch.bfloat16).contiguous()
    dx = torch.empty_like(x)
    lib = _lib.load()
    cs = torch.empty(N, dtype=torch.float32, device=x.device) if want_colsum else None
    ws = _workspace(lib.sb200_gelu_bwd_workspace_bytes(R, N), x.device) if want_colsum else None
    with torch.cuda.device(x.device):
        code = lib.sb200_gelu_bwd(_ptr(x), _ptr(dy), R, N, _ptr(dx), _ptr(cs), _ptr(ws), 0 if ws is None else ws.numel(),
                                  _stream())
    _lib.check(code, "sb200_gelu_bwd")
    return dx, cs


class LinearGeluFunction(torch.autograd.Function):
    """gelu(x W^T + b) for bf16 operands (transformers BertIntermediate, BertPredictionHeadTransform.dense + act):
    cuBLAS GEMMs, GELU forward kernel, and one backward kernel that produces the pre-activation gradient and the
    bias gradient together."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        w = half_weight(weight, x.dtype)
        pre = torch.nn.functional.linear(x, w, None if bias is None else half_weight(bias, x.dtype))
        ctx.save_for_backward(x, w, pre)
        ctx.has_bias = bias is not None
        ctx.w_dtype = weight.dtype
        return gelu_forward(pre)

    @staticmethod
    def backward(ctx, dy):
        x, weight, pre = ctx.saved_tensors
        N = pre.shape[-1]
        want_db = ctx.has_bias and ctx.needs_input_grad[2]
        dpre, db = gelu_backward(pre.view(-1, N), dy.reshape(-1, N), want_colsum=want_db)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = (dpre @ weight).view(x.shape)
        if ctx.needs_input_grad[1]:
            dw = _weight_grad(dpre, x.reshape(-1, x.shape[-1]), ctx.w_dtype)
        return dx, dw, db


def linear_gelu(x, weight, bias):
    """gelu(linear(x)) under bf16 autocast (or with bf16 operands); raises otherwise (bf16-only kernels)."""
    if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
        with torch.autocast("cuda", enabled=False):
            return LinearGeluFunction.apply(x.to(torch.bfloat16), weight, bias)
    if x.dtype != torch.bfloat16:
        raise TypeError("linear_gelu: bf16 operands or bf16 autocast required")
    return LinearGeluFunction.apply(x, weight, bias)


def linear(x, weight, bias):
    """torch.nn.functional.linear semantics, including autocast: under autocast the operands are cast to the active
    autocast dtype (bf16 or fp16) exactly like the stock op, and the matmuls run outside the autocast region."""
    if torch.is_autocast_enabled("cuda"):
        dt = torch.get_autocast_dtype("cuda")
        with torch.autocast("cuda", enabled=False):
            return LinearFunction.apply(x.to(dt), weight, bias)
    return LinearFunction.apply(x, weight, bias)
