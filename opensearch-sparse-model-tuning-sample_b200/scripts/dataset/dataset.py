"""Training datasets of the reference's ``scripts/dataset/dataset.py`` that feed the hot path: ``kd`` rows
(query, docs, teacher scores; reference :150-213) and ``posnegs`` rows (query, positive, hard negatives; :329-352),
plus ``load_dataset`` (:452-467). The BEIR / MIRACL evaluation corpora and the DynamoDB ``kd-ids`` variant belong to
the OpenSearch side of the reference and are out of scope.

A dataset directory written by ``datasets.Dataset.save_to_disk`` is read with the ``datasets`` package like upstream;
a ``.jsonl`` / ``.json`` file with the same columns is accepted as well (offline use).
"""
import json
import logging
import os
import random

from torch.utils.data import Dataset

logger = logging.getLogger(__name__)


def _read_rows(path):
    if os.path.isdir(path):
        import datasets
        return datasets.Dataset.load_from_disk(path)
    with open(path) as f:
        if path.endswith(".jsonl"):
            return [json.loads(line) for line in f if line.strip()]
        return json.load(f)


def _columns(rows):
    names = getattr(rows, "column_names", None)
    return set(names) if names is not None else (set(rows[0]) if len(rows) else set())


def _partial_shuffle(order, swap_times, rng):
    """swap_times random transpositions (a full shuffle once that is at least half the list) -- reference :23-42."""
    order = list(order)
    n = len(order)
    if swap_times <= 0 or n < 2:
        return order
    if swap_times >= n // 2:
        rng.shuffle(order)
        return order
    for _ in range(int(swap_times)):
        i, j = rng.randrange(n), rng.randrange(n)
        order[i], order[j] = order[j], order[i]
    return order


class KnowledgeDistillDataset(Dataset):
    """Rows {"query", "docs": [...], "scores": [...] (optional), "first_rank" (optional)}. Every row is cut into
    len(docs) // sample_num groups; group i takes the docs i, i+step, i+2*step, ... of the (partially shuffled) list, so a
    group spans the teacher's ranking from top to bottom. Items are (query, docs, scores)."""

    def __init__(self, all_data, sample_num=2, swap_times=0, first_rank_thresh=1000, score_scale=1.0, seed=0, **unused):
        assert sample_num >= 2
        cols = _columns(all_data)
        if "first_rank" in cols:
            keep = [i for i, ex in enumerate(all_data) if 0 <= ex.get("first_rank", 1) <= first_rank_thresh]
            all_data = all_data.select(keep) if hasattr(all_data, "select") else [all_data[i] for i in keep]
        self.all_data = all_data
        self.has_scores = "scores" in cols
        self.score_scale = score_scale
        rng = random.Random(seed)
        self.idxs = []
        for ex_idx, ex in enumerate(all_data):
            order = _partial_shuffle(range(len(ex["docs"])), swap_times, rng)
            step = len(order) // sample_num
            self.idxs.extend((ex_idx, [order[k * step + i] for k in range(sample_num)]) for i in range(step))
        logger.info("KnowledgeDistillDataset: %d rows -> %d samples", len(all_data), len(self.idxs))

    def __len__(self):
        return len(self.idxs)

    def __getitem__(self, idx):
        ex_idx, picks = self.idxs[idx]
        ex = self.all_data[ex_idx]
        docs = [ex["docs"][i] for i in picks]
        scores = [ex["scores"][i] * self.score_scale for i in picks] if self.has_scores else [None] * len(picks)
        return ex["query"], docs, scores


class PosNegsDataset(Dataset):
    """Rows {"query", "pos", "negs": [...]}: one sample per full window of sample_num hard negatives. Items are
    [query, pos, negs_window]."""

    def __init__(self, data, sample_num=3, **unused):
        assert sample_num >= 1
        self.data = []
        for row in data:
            negs = row.get("negs", []) or []
            for i in range(0, len(negs) - sample_num + 1, sample_num):
                self.data.append([row["query"], row["pos"], list(negs[i:i + sample_num])])
        logger.info("PosNegsDataset: %d rows -> %d samples", len(data), len(self.data))

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        return self.data[idx]


DATASET_CLS_MAP = {"kd": KnowledgeDistillDataset, "posnegs": PosNegsDataset}


def load_dataset(path, cls, swap_times=0, sample_num_one_query=2, first_rank_thresh=1000, score_scale=1.0):
    if cls not in DATASET_CLS_MAP:
        raise NotImplementedError(f"data_type={cls!r}: supported here are {sorted(DATASET_CLS_MAP)} (+ 'synthetic', 'tensors'); "
                                  "'kd-ids' needs the reference's DynamoDB embedding service, which is out of scope")
    logger.info("load dataset from %s. dataset cls: %s", path, DATASET_CLS_MAP[cls].__name__)
    return DATASET_CLS_MAP[cls](_read_rows(path), sample_num=sample_num_one_query, swap_times=swap_times,
                                first_rank_thresh=first_rank_thresh, score_scale=score_scale)
