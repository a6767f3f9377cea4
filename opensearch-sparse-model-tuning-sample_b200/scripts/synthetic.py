"""Offline stand-ins used by bench.py, smoke() and the tests: random-init backbones of the reference model shapes,
a tokenizer object with the BERT-uncased special-token layout, and the seeded synthetic token batches of
SURVEY.md section 8(d). No pretrained weights or text corpora are reachable in this environment.
"""
import torch

VOCAB_SIZE = 30522
SPECIAL = {"pad_token": ("[PAD]", 0), "unk_token": ("[UNK]", 100), "cls_token": ("[CLS]", 101),
           "sep_token": ("[SEP]", 102), "mask_token": ("[MASK]", 103)}

MODEL_SHAPES = {
    # opensearch-neural-sparse-encoding-doc-v2-mini: 6 layers x 384 (22.7M parameters)
    "mini": dict(hidden_size=384, num_hidden_layers=6, num_attention_heads=12, intermediate_size=1536),
    # BERT-base (co-condenser-marco, opensearch-neural-sparse-encoding-v1)
    "base": dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072),
    # kd-ensemble dense teacher stand-in (Alibaba-NLP/gte-large-en-v1.5 scale: 24 layers x 1024)
    "large": dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096),
    "tiny": dict(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128),
}


class SyntheticTokenizer:
    """Carries only what SparseModel reads: vocab, special_tokens_map, token<->id conversion."""

    def __init__(self, vocab_size=VOCAB_SIZE):
        self.vocab = {f"tok{i}": i for i in range(vocab_size)}
        self.special_tokens_map = {}
        for key, (tok, idx) in SPECIAL.items():
            if idx < vocab_size:
                self.vocab.pop(f"tok{idx}")
                self.vocab[tok] = idx
                self.special_tokens_map[key] = tok
        self._inv = {i: t for t, i in self.vocab.items()}

    def _convert_token_to_id_with_added_voc(self, token):
        return self.vocab[token]

    def __call__(self, texts, padding=True, truncation=True, max_length=512, return_tensors="pt",
                 return_token_type_ids=False, **unused):
        """Whitespace "tokenisation" with the HF call signature (offline stand-in): every word is hashed (crc32) to a
        non-special id; [CLS] ... [SEP], right-padded with [PAD] to the longest text of the batch."""
        import zlib
        V = len(self.vocab)
        low = min(1000, V // 2)
        rows = []
        for t in texts:
            ids = [low + zlib.crc32(w.encode()) % (V - low) for w in t.split()][:max(0, max_length - 2)]
            rows.append([min(101, V - 1)] + ids + [min(102, V - 1)])
        width = max(len(r) for r in rows)
        out_ids = torch.zeros(len(rows), width, dtype=torch.long)
        mask = torch.zeros(len(rows), width, dtype=torch.long)
        for i, r in enumerate(rows):
            out_ids[i, :len(r)] = torch.tensor(r)
            mask[i, :len(r)] = 1
        return {"input_ids": out_ids, "attention_mask": mask}

    def _convert_id_to_token(self, idx):
        return self._inv[int(idx)]

    def save_pretrained(self, output_dir):
        import json
        import os
        with open(os.path.join(output_dir, "synthetic_vocab.json"), "w") as f:
            json.dump({"size": len(self.vocab)}, f)


def build_backbone(shape="mini", vocab_size=VOCAB_SIZE, seed=0, max_position_embeddings=512, dropout=0.1):
    import transformers
    cfg = transformers.BertConfig(vocab_size=vocab_size, max_position_embeddings=max_position_embeddings,
                                  hidden_dropout_prob=dropout, attention_probs_dropout_prob=dropout,
                                  **MODEL_SHAPES[shape])
    torch.manual_seed(seed)
    return transformers.BertForMaskedLM(cfg)


def build_sparse_model(shape="mini", idf_vector=None, use_l0=False, vocab_size=VOCAB_SIZE, seed=0, bias_shift=0.0,
                       prune_ratio=None, idf_requires_grad=False, dropout=0.1, fuse_body=True, unpad_capacity=None,
                       attention="auto"):
    """SparseModel over a random-init backbone. bias_shift < 0 gives the "trained-like" activation regime."""
    from .model.sparse_encoders import SparseModel
    backbone = build_backbone(shape, vocab_size, seed, dropout=dropout)
    model = SparseModel(None, backbone=backbone, tokenizer=SyntheticTokenizer(vocab_size), use_l0=use_l0,
                        prune_ratio=prune_ratio, idf_requires_grad=idf_requires_grad, fuse_body=fuse_body,
                        unpad_capacity=unpad_capacity, attention=attention)
    if idf_vector is not None:
        with torch.no_grad():
            model.idf_vector.copy_(idf_vector)
    if bias_shift != 0.0:
        with torch.no_grad():
            backbone.cls.predictions.decoder.bias.add_(bias_shift)
    return model


def token_batch(batch, seq_len, seed, vocab_size=VOCAB_SIZE, full_length=False, device="cpu"):
    """ids uniform in [1000, V), [CLS] first, [SEP] last real token, lengths uniform in [L/2, L], [PAD]=0 tail."""
    g = torch.Generator().manual_seed(seed)
    low = min(1000, vocab_size // 2)
    ids = torch.randint(low, vocab_size, (batch, seq_len), generator=g)
    if full_length or seq_len < 4:
        lens = torch.full((batch,), seq_len)
    else:
        lens = torch.randint(seq_len // 2, seq_len + 1, (batch,), generator=g)
    pos = torch.arange(seq_len)[None, :]
    mask = (pos < lens[:, None]).long()
    ids[:, 0] = min(101, vocab_size - 1)
    ids[torch.arange(batch), lens - 1] = min(102, vocab_size - 1)
    ids = ids * mask
    return {"input_ids": ids.to(device), "attention_mask": mask.to(device)}


def train_batch(n_queries, docs_per_query, doc_len, query_len=32, seed=1234, vocab_size=VOCAB_SIZE, device="cpu",
                with_scores=None, n_feature_sets=1):
    """The dict layout compute_loss expects (reference collator.py:23-57, 146-177): lists, element 0 = student,
    elements 1.. = the same texts tokenised for each kd-ensemble teacher (here: copies of the student's ids)."""
    q = token_batch(n_queries, query_len, seed, vocab_size, device=device)
    d = token_batch(n_queries * docs_per_query, doc_len, seed + 1, vocab_size, device=device)
    batch = {"query": [q] + [{k: v.clone() for k, v in q.items()} for _ in range(n_feature_sets - 1)],
             "docs": [d] + [{k: v.clone() for k, v in d.items()} for _ in range(n_feature_sets - 1)]}
    if with_scores is not None:
        g = torch.Generator().manual_seed(seed + 11)
        batch["scores"] = (torch.randn(n_queries, with_scores, generator=g) * 3).to(device)
    return batch
