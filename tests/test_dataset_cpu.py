"""Host-side data path of train_ir.py (reference scripts/dataset/dataset.py:150-213, 329-352 and collator.py:8-57, 146-177):
sample construction, the batch dict layout compute_loss consumes (docs query-major, positive first; element 0 of each
list = student tokenisation, 1.. = teachers), and train_ir's dataset dispatch. CPU only."""
import json
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparse_b200  # noqa: F401,E402
from sparse_b200.scripts.dataset.collator import COLLATOR_CLS_MAP  # noqa: E402
from sparse_b200.scripts.dataset.dataset import (DATASET_CLS_MAP, KnowledgeDistillDataset, PosNegsDataset,  # noqa: E402
                                                   load_dataset)


class WordTokenizer:
    """Whitespace tokenizer with the HF call signature the collators use."""

    def __init__(self, offset=0):
        self.offset = offset

    def __call__(self, texts, padding=True, truncation=True, max_length=8, return_tensors="pt", return_token_type_ids=False):
        rows = [[101] + [self.offset + 1000 + len(w) for w in t.split()][:max_length - 2] + [102] for t in texts]
        width = max(len(r) for r in rows)
        ids = torch.zeros(len(rows), width, dtype=torch.long)
        mask = torch.zeros(len(rows), width, dtype=torch.long)
        for i, r in enumerate(rows):
            ids[i, :len(r)] = torch.tensor(r)
            mask[i, :len(r)] = 1
        return {"input_ids": ids, "attention_mask": mask}


def test_kd_dataset_groups_span_the_ranking():
    rows = [{"query": "q0", "docs": [f"d{i}" for i in range(6)], "scores": [6.0, 5, 4, 3, 2, 1]},
            {"query": "q1", "docs": ["a", "b", "c"], "scores": [3.0, 2, 1]}]
    ds = KnowledgeDistillDataset(rows, sample_num=2, score_scale=2.0)
    assert len(ds) == 3 + 1                                   # 6 // 2 groups, 3 // 2 groups
    q, docs, scores = ds[0]
    assert (q, docs, scores) == ("q0", ["d0", "d3"], [12.0, 6.0])   # group i = docs i, i + step
    assert ds[2][1] == ["d2", "d5"] and ds[3][1] == ["a", "b"]
    # first_rank filter and score-less rows
    ds2 = KnowledgeDistillDataset([{"query": "q", "docs": ["x", "y"], "first_rank": 5000},
                                   {"query": "r", "docs": ["x", "y"], "first_rank": 3}], sample_num=2, first_rank_thresh=1000)
    assert len(ds2) == 1 and ds2[0] == ("r", ["x", "y"], [None, None])


def test_posnegs_dataset_windows():
    rows = [{"query": "q", "pos": "p", "negs": ["n0", "n1", "n2", "n3", "n4"]}, {"query": "r", "pos": "s", "negs": ["m0"]}]
    ds = PosNegsDataset(rows, sample_num=2)
    assert [ds[i] for i in range(len(ds))] == [["q", "p", ["n0", "n1"]], ["q", "p", ["n2", "n3"]]]


def test_collators_layout():
    kd = COLLATOR_CLS_MAP["kd"](WordTokenizer(), 8, [WordTokenizer(offset=50)])
    batch = kd([("q zero", ["aa bb cc", "d"], [2.0, 1.0]), ("q one two", ["e f", "g h i j k l m n"], [0.5, 0.25])])
    assert len(batch["query"]) == 2 and len(batch["docs"]) == 2            # student + one teacher tokenisation
    assert batch["docs"][0]["input_ids"].shape[0] == 4 and batch["query"][0]["input_ids"].shape[0] == 2
    assert batch["docs"][0]["input_ids"].shape[1] == 8                      # truncated at max_length
    assert torch.equal(batch["scores"], torch.tensor([[2.0, 1.0], [0.5, 0.25]]))
    assert int(batch["docs"][1]["input_ids"][0, 1]) == int(batch["docs"][0]["input_ids"][0, 1]) + 50
    pn = COLLATOR_CLS_MAP["posnegs"](WordTokenizer(), 16)
    b2 = pn([["q", "pos one", ["neg a", "neg bb"]], ["r", "pos two three", ["neg c", "neg d"]]])
    ids = b2["docs"][0]["input_ids"]
    assert ids.shape[0] == 6 and "scores" not in b2
    # query-major, positive first: rows 0 and 3 are the positives ("pos one" -> 2 words, "pos two three" -> 3 words)
    assert int(b2["docs"][0]["attention_mask"][0].sum()) == 4 and int(b2["docs"][0]["attention_mask"][3].sum()) == 5


def test_load_dataset_from_jsonl_and_train_ir_dispatch(tmp_path):
    path = tmp_path / "train.jsonl"
    with open(path, "w") as f:
        for i in range(5):
            f.write(json.dumps({"query": f"q {i}", "pos": f"p {i}", "negs": [f"n {i} {k}" for k in range(4)]}) + "\n")
    ds = load_dataset(str(path), "posnegs", sample_num_one_query=2)
    assert len(ds) == 10 and set(DATASET_CLS_MAP) == {"kd", "posnegs"}
    with pytest.raises(NotImplementedError):
        load_dataset(str(path), "kd-ids")
    from types import SimpleNamespace
    from sparse_b200 import train_ir
    data_args = SimpleNamespace(data_type="posnegs", train_file=str(path), swap_times=0, sample_num_one_query=2,
                                first_rank_thresh=1000, max_seq_length=16, kd_ensemble_teacher_kwargs={},
                                loss_types=["infonce"])
    targs = SimpleNamespace(per_device_train_batch_size=3, max_steps=4)
    model = SimpleNamespace(tokenizer=WordTokenizer(), vocab_size=30522)
    dataset, collator, bs = train_ir.build_dataset(data_args, targs, model, 0)
    assert bs == 3 and len(dataset) == 10
    batch = collator([dataset[i] for i in range(3)])
    assert batch["docs"][0]["input_ids"].shape[0] == 9 and batch["query"][0]["input_ids"].shape[0] == 3
    # synthetic data keeps yielding pre-collated batches (DataLoader batch size 1)
    data_args.data_type = "synthetic"
    dataset, collator, bs = train_ir.build_dataset(data_args, targs, model, 0)
    assert bs == 1 and collator([dataset[0]]) is dataset[0]
