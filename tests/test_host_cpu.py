"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol of include/sparse_b200.h, the
product path refuses to run without CUDA (no fallback), configs parse with the reference's keys."""
import os
import re

import pytest
import torch
import yaml

import sparse_b200
from sparse_b200 import _lib, ops
from conftest import ROOT


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "sparse_b200.h")).read()
    declared = set(re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sb200_abi_version() == 3
    assert lib.sb200_head_fwd_workspace_bytes(160, 256) > 0
    assert lib.sb200_head_bwd_workspace_bytes(160, 256, 384, 30522) >= 160 * 30522 * 8


def test_ctypes_prototypes_match_the_header():
    """Every ctypes prototype of _lib.py has the argument count and the scalar/pointer kinds of its declaration in
    include/sparse_b200.h (a mismatch would corrupt the stack silently, and no CPU test calls the kernels)."""
    import ctypes
    header = open(os.path.join(ROOT, "include", "sparse_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = dict(re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S))
    assert set(protos) == set(_lib._PROTOTYPES)
    for name, params in protos.items():
        params = " ".join(params.split())
        args = [] if params in ("", "void") else [a.strip() for a in params.split(",")]
        restype, argtypes = _lib._PROTOTYPES[name]
        assert len(args) == len(argtypes), (name, args, argtypes)
        for a, t in zip(args, argtypes):
            is_ptr = "*" in a or a.startswith("sb200_stream_t")
            if is_ptr:
                assert t is ctypes.c_void_p, (name, a, t)
            elif a.startswith("float"):
                assert t is ctypes.c_float, (name, a, t)
            elif a.startswith("size_t"):
                assert t is ctypes.c_size_t, (name, a, t)
            elif a.startswith("int ") or a.startswith("int32_t"):
                assert t is ctypes.c_int, (name, a, t)
            else:
                raise AssertionError(f"{name}: unclassified parameter {a!r}")


def test_argument_errors_are_reported_not_crashes():
    lib = _lib.load()
    code = lib.sb200_head_fwd(0, 0, 0, 0, 8, 1, 1, 8, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0)
    assert code == 1 and b"null" in lib.sb200_last_error()
    code = lib.sb200_rank_loss(7, 1, 0, 1, 1, 1, 0, 1.0, 1, 0, 0, 0, 0)
    assert code == 1 and b"mode" in lib.sb200_last_error()


def test_no_cpu_fallback():
    h = torch.zeros(1, 4, 8, dtype=torch.bfloat16)
    w = torch.zeros(16, 8, dtype=torch.bfloat16)
    with pytest.raises(_lib.SparseB200Error):
        ops.head_forward(h, w, None, torch.ones(1, 4, dtype=torch.long))
    with pytest.raises(_lib.SparseB200Error):
        ops.flops_value(torch.zeros(4, 8), 1)
    with pytest.raises(_lib.SparseB200Error):
        ops.scores(torch.zeros(2, 8), torch.zeros(4, 8), True)
    with pytest.raises(_lib.SparseB200Error):
        ops.varlen_attention(torch.zeros(4, 3, 2, 32, dtype=torch.bfloat16), torch.tensor([0, 4], dtype=torch.int32), 4)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "opensearch-sparse-model-tuning-sample_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                assert "oracle" not in src.replace("oracle/", ""), os.path.join(base, f)


def test_reference_configs_parse():
    from sparse_b200.scripts.args import parse_dict
    cfg_dir = os.path.join(ROOT, "opensearch-sparse-model-tuning-sample_b200", "configs")
    for name in ("config_infonce.yaml", "config_kd.yaml", "config_l0.yaml"):
        m, d, t = parse_dict(yaml.safe_load(open(os.path.join(cfg_dir, name))))
        assert m.inf_free is True and t.max_steps > 0 and d.loss_types
    m, d, t = parse_dict(yaml.safe_load(open(os.path.join(cfg_dir, "config_l0.yaml"))))
    assert d.flops_threshold == 150 and m.use_l0 is True and d.loss_types == ["kldiv"]


def test_trainer_lambda_schedule_matches_reference_values(golden):
    from sparse_b200.scripts.args import DataTrainingArguments, ModelArguments
    from sparse_b200.scripts.train.trainer import SparseModelTrainer

    class Dummy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))
    tr = SparseModelTrainer(ModelArguments(), DataTrainingArguments(), [], model=Dummy())
    for step, expect in golden["lambda"]:
        tr.state.global_step = step
        assert tr.get_lambda(0.05, 200) == pytest.approx(expect, rel=1e-12)


def test_synthetic_batches_layout():
    from sparse_b200.scripts import synthetic
    b = synthetic.train_batch(4, 3, 64, query_len=16)
    docs = b["docs"][0]
    assert docs["input_ids"].shape == (12, 64) and docs["attention_mask"].shape == (12, 64)
    assert (docs["input_ids"][:, 0] == 101).all()
    lens = docs["attention_mask"].sum(1)
    assert (docs["input_ids"][torch.arange(12), lens - 1] == 102).all()
    assert ((docs["input_ids"] == 0) == (docs["attention_mask"] == 0)).all()
    tok = synthetic.SyntheticTokenizer()
    assert sorted(tok.vocab[t] for t in tok.special_tokens_map.values()) == [0, 100, 101, 102, 103]


def test_packed_body_plan_and_repad_cpu():
    """Host-side logic of the padding-free body: device-free placement plan (holes, empty rows, overflow) and the
    scatter back to [B, L, H] with its gradient."""
    from sparse_b200.scripts.model.packed_body import PackedBertBody
    body = PackedBertBody.__new__(PackedBertBody)
    body.capacity, body.overflow_count = 1.0, None
    mask = torch.tensor([[1, 1, 1, 0], [1, 0, 1, 1], [0, 0, 0, 0], [1, 1, 1, 1]])
    t_cap, dest, src_of, row_valid, cu = body._plan(mask)
    assert t_cap == 16 and int(body.overflow_count) == 0
    kept = mask.reshape(-1).nonzero().flatten()
    assert dest[kept].tolist() == list(range(10)) and set(dest[mask.reshape(-1) == 0].tolist()) == {16}
    assert src_of[:10].tolist() == kept.tolist() and row_valid.tolist() == [True] * 10 + [False] * 6
    assert cu.tolist() == [0, 3, 6, 6, 10, 14, 16, 16, 16]          # real sequences, then dummy ones over the filler rows
    packed = torch.arange(16.0).unsqueeze(1).repeat(1, 2).requires_grad_(True)
    plan = (dest.clamp_max(t_cap - 1), src_of, row_valid, (4, 4))
    padded = PackedBertBody.repad(packed, plan)
    assert padded.shape == (4, 4, 2)
    assert padded.reshape(16, 2)[kept, 0].tolist() == list(range(10))
    w = torch.arange(1.0, 17.0).unsqueeze(1).repeat(1, 2)
    (padded.reshape(16, 2) * w).sum().backward()
    assert packed.grad[:10, 0].tolist() == (kept + 1).float().tolist()   # gradient of the padded position it fed
    assert packed.grad[10:].abs().sum() == 0                              # filler rows get none
    # capacity below the token count: counted, tokens past the capacity are dropped into the dummy slot
    body.capacity, body.overflow_count = 0.5, None
    t_cap, dest, src_of, row_valid, cu = body._plan(mask)
    assert t_cap == 8 and int(body.overflow_count) == 1
    assert dest[kept].tolist() == list(range(8)) + [8, 8] and int(cu.max()) == 8 and cu.tolist() == sorted(cu.tolist())
    body._plan(mask[:2])
    assert int(body.overflow_count) == 1                                  # 6 tokens fit the 8 rows: counter unchanged


def test_attention_selection_of_the_packed_body():
    """`attention` (auto / own / flash) is validated when the body is built, is a ModelArguments key, and the packed
    body hands the number of real sequences to the attention op (the filler sequences behind them are only
    zero-filled)."""
    import inspect
    from sparse_b200.scripts import synthetic
    from sparse_b200.scripts.args import ModelArguments, parse_dict
    from sparse_b200.scripts.model.packed_body import PackedBertBody
    bert = synthetic.build_backbone("tiny", 200, 0).bert                 # 4 heads x 16: outside the own kernels
    with pytest.raises(ValueError):
        PackedBertBody(bert, 1.0, attention="bogus")
    with pytest.raises(RuntimeError):
        PackedBertBody(bert, 1.0, attention="own")
    assert ops.attn_supported(32, 256) and ops.attn_supported(64, 512)
    assert not ops.attn_supported(16, 64) and not ops.attn_supported(32, 2048)
    assert ModelArguments().attention == "auto"
    m, _, _ = parse_dict({"attention": "flash", "unpad_capacity": 0.9})
    assert m.attention == "flash" and m.unpad_capacity == 0.9
    assert "live_sequences=B" in inspect.getsource(PackedBertBody.__call__)


def test_trainer_refuses_overflowed_packed_batches():
    """A batch that did not fit unpad_capacity is reported at logging cadence, never trained on silently."""
    from types import SimpleNamespace
    from sparse_b200.scripts.train.trainer import SparseModelTrainer
    t = SparseModelTrainer.__new__(SparseModelTrainer)
    t.model_wrapper = SimpleNamespace(sparse_model=SimpleNamespace(unpad_overflows=lambda: 0))
    t._check_unpad()
    t.model_wrapper.sparse_model.unpad_overflows = lambda: 2
    with pytest.raises(RuntimeError, match="unpad_capacity"):
        t._check_unpad()
