"""Collators of the reference's ``scripts/dataset/collator.py`` (``kd`` :8-57, ``posnegs`` :146-177): texts -> the
batch dict ``compute_loss`` consumes, {"query": [features per tokenizer], "docs": [features per tokenizer],
"scores": tensor (kd only)}, docs query-major with the positive first. Element 0 of each list is the student's
tokenisation, elements 1.. the kd-ensemble teachers' (``teacher_tokenizer_ids``).

``PrefetchLoader`` is the B200-side addition: batches are pinned and copied host->device on
a side stream one step ahead, so the copy of batch i+1 overlaps the compute of batch i.
"""
import itertools
import logging

import torch

logger = logging.getLogger(__name__)


def _load_tokenizers(tokenizer, teacher_tokenizer_ids):
    import transformers
    extra = [t if not isinstance(t, str) else transformers.AutoTokenizer.from_pretrained(t) for t in teacher_tokenizer_ids]
    logger.info("total tokenizers %d", 1 + len(extra))
    return [tokenizer] + extra


class _TextCollator:
    def __init__(self, tokenizer, max_length=512, teacher_tokenizer_ids=(), **unused):
        self.tokenizer = tokenizer
        self.max_length = max_length
        self.tokenizers = _load_tokenizers(tokenizer, list(teacher_tokenizer_ids))
        if unused:
            logger.info("unused args: %s", unused)

    def _encode_all(self, queries, docs):
        kw = dict(padding=True, truncation=True, max_length=self.max_length, return_tensors="pt",
                  return_token_type_ids=False)
        queries, docs = list(queries), list(docs)      # every tokenizer sees the same texts (docs may be an iterator)
        return {"query": [tok(queries, **kw) for tok in self.tokenizers],
                "docs": [tok(docs, **kw) for tok in self.tokenizers]}


class KnowledgeDistillDataCollator(_TextCollator):
    def __call__(self, batch):
        queries, docs, scores = zip(*batch)
        assert len(docs) == len(scores)
        result = self._encode_all(queries, itertools.chain.from_iterable(docs))
        if scores[0][0] is not None:
            result["scores"] = torch.tensor(scores)
        return result


class PosNegsDataCollator(_TextCollator):
    def __call__(self, batch):
        queries, positives, negatives = zip(*batch)
        assert len(queries) == len(positives)
        docs = []
        for pos, negs in zip(positives, negatives):
            docs.append(pos)
            docs.extend(negs)
        return self._encode_all(queries, docs)


COLLATOR_CLS_MAP = {"kd": KnowledgeDistillDataCollator, "posnegs": PosNegsDataCollator}


def _map_tensors(obj, fn):
    if torch.is_tensor(obj):
        return fn(obj)
    if hasattr(obj, "items"):
        return {k: _map_tensors(v, fn) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_map_tensors(v, fn) for v in obj]
    return obj


class PrefetchLoader:
    """Wraps a DataLoader: yields device-resident batches; the H2D copy of the next batch runs on its own stream."""

    def __init__(self, loader, device):
        self.loader = loader
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.sampler = getattr(loader, "sampler", None)

    def __len__(self):
        return len(self.loader)

    def _stage(self, batch):
        with torch.cuda.stream(self.stream):
            return _map_tensors(batch, lambda t: (t if t.is_pinned() else t.pin_memory()).to(self.device, non_blocking=True))

    def __iter__(self):
        staged = None
        for batch in self.loader:
            nxt = self._stage(batch)
            if staged is not None:
                yield staged
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
            # the consumer's stream now owns the tensors: tell the caching allocator
            _map_tensors(nxt, lambda t: t.record_stream(torch.cuda.current_stream(self.device)) or t)
            staged = nxt
        if staged is not None:
            yield staged
