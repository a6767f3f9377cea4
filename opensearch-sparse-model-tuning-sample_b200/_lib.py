"""ctypes binding of libsparse_b200.so (C ABI in include/sparse_b200.h).

The product path has no CPU fallback: if the shared library is missing or a call fails, this module raises.
Build the library with ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C csrc``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsparse_b200.so")

SB200_OK = 0
HEAD_L0 = 1
HEAD_FP16 = 2
LOSS_INFONCE, LOSS_KLDIV, LOSS_MARGINMSE = 0, 1, 2

_c_int = ctypes.c_int
_c_f = ctypes.c_float
_vp = ctypes.c_void_p
_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/sparse_b200.h one to one
_PROTOTYPES = {
    "sb200_abi_version": (_c_int, []),
    "sb200_last_error": (ctypes.c_char_p, []),
    "sb200_launch_count": (ctypes.c_ulonglong, []),
    "sb200_head_fwd_workspace_bytes": (_sz, [_c_int, _c_int]),
    "sb200_head_fwd": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp,
                                _vp, _c_int, _vp, _sz, _vp]),
    "sb200_head_bwd_workspace_bytes": (_sz, [_c_int, _c_int, _c_int, _c_int]),
    "sb200_head_bwd": (_c_int, [_vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp,
                                _sz, _vp]),
    "sb200_head_packed_supported": (_c_int, [_c_int, _c_int]),
    "sb200_head_fwd_packed": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp,
                                       _vp, _c_int, _vp, _sz, _vp]),
    "sb200_head_bwd_packed": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp,
                                       _vp, _vp, _vp, _sz, _vp]),
    "sb200_prune_rows": (_c_int, [_vp, _c_int, _c_int, _c_f, _vp]),
    "sb200_idf_query": (_c_int, [_vp, _c_int, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp]),
    "sb200_idf_query_bwd": (_c_int, [_vp, _vp, _c_int, _c_int, _vp, _vp]),
    "sb200_flops_workspace_bytes": (_sz, [_c_int, _c_int, _c_int]),
    "sb200_flops_fwd": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sb200_flops_bwd": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "sb200_scores_workspace_bytes": (_sz, [_c_int, _c_int, _c_int, _c_int]),
    "sb200_scores_fwd": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _sz, _vp]),
    "sb200_scores_bwd": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                  _c_int, _vp, _vp, _vp, _sz, _vp]),
    "sb200_rank_loss_workspace_bytes": (_sz, [_c_int]),
    "sb200_rank_loss": (_c_int, [_c_int, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_f, _vp, _vp, _vp, _sz, _vp]),
    "sb200_score_loss_fwd": (_c_int, [_c_int, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f, _c_int, _vp,
                                      _vp, _vp, _vp, _sz, _vp]),
    "sb200_compact_workspace_bytes": (_sz, [_c_int, _c_int]),
    "sb200_compact_rows": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp, _vp, _vp, _c_int, _vp, _vp, _sz, _vp]),
    "sb200_minmax_accumulate": (_c_int, [_vp, _c_int, _c_int, _c_f, _c_int, _vp, _vp]),
    "sb200_layer_norm_supported": (_c_int, [_c_int]),
    "sb200_layer_norm_fwd": (_c_int, [_vp, _c_int, _vp, _vp, _c_int, _c_int, _c_f, _vp, _vp, _vp, _vp]),
    "sb200_layer_norm_bwd_workspace_bytes": (_sz, [_c_int, _c_int]),
    "sb200_layer_norm_bwd": (_c_int, [_vp, _vp, _c_int, _vp, _vp, _vp, _c_int, _c_int, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sb200_add_layer_norm_fwd": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_f, _vp, _c_f, _vp, _vp, _vp, _vp, _vp]),
    "sb200_add_layer_norm_bwd": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _vp, _c_f, _vp, _vp, _vp,
                                          _vp, _vp, _sz, _vp]),
    "sb200_gelu_fwd": (_c_int, [_vp, _sz, _vp, _vp]),
    "sb200_gelu_bwd_workspace_bytes": (_sz, [_c_int, _c_int]),
    "sb200_gelu_bwd": (_c_int, [_vp, _vp, _c_int, _c_int, _vp, _vp, _vp, _sz, _vp]),
    "sb200_embed_sum_fwd": (_c_int, [_vp] * 6 + [_c_int] * 5 + [_vp, _vp]),
    "sb200_embed_sum_bwd": (_c_int, [_vp] * 4 + [_c_int] * 6 + [_vp] * 4),
    "sb200_colsum_supported": (_c_int, [_c_int]),
    "sb200_colsum_workspace_bytes": (_sz, [_c_int, _c_int]),
    "sb200_colsum": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp, _vp, _sz, _vp]),
    "sb200_attn_supported": (_c_int, [_c_int, _c_int]),
    "sb200_attn_fwd": (_c_int, [_vp] * 3 + [_sz, _vp] + [_c_int] * 6 + [_c_f, _c_f, _vp, _c_int, _vp, _vp, _vp]),
    "sb200_attn_bwd": (_c_int, [_vp] * 3 + [_sz] + [_vp] * 4 + [_c_int] * 6 + [_c_f, _c_f, _vp, _c_int] + [_vp] * 3
                       + [_sz, _vp, _vp]),
    "sb200_attn_dropout_mask": (_c_int, [_vp] + [_c_int] * 4 + [_c_f, _vp, _c_int, _vp, _vp]),
    "sb200_peer_alloc": (_c_int, [_sz, _vp]),
    "sb200_peer_free": (_c_int, [_vp]),
    "sb200_peer_export": (_c_int, [_vp, _vp]),
    "sb200_peer_import": (_c_int, [_vp, _vp]),
    "sb200_peer_close": (_c_int, [_vp]),
    "sb200_peer_allgather": (_c_int, [_vp, _sz, _c_int, _c_int, _vp, _sz, _vp]),
    "sb200_peer_signal": (_c_int, [_vp, _c_int, _c_int, _vp, _sz, _vp]),
    "sb200_peer_wait": (_c_int, [_vp, _vp, _c_int, _vp]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None


class SparseB200Error(RuntimeError):
    pass


def load():
    """Loads the shared library once; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SparseB200Error(
            f"{LIB_PATH} is missing: the sm_100a CUDA library is not built and there is no fallback path. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` at the repository root."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise SparseB200Error(f"{LIB_PATH} does not export {name}; rebuild the library") from None
        fn.restype = res
        fn.argtypes = args
    if lib.sb200_abi_version() != 3:
        raise SparseB200Error("libsparse_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(code, what):
    if code != SB200_OK:
        msg = load().sb200_last_error()
        raise SparseB200Error(f"{what} failed (code {code}): {msg.decode() if msg else '?'}")


def launch_count():
    return int(load().sb200_launch_count())
