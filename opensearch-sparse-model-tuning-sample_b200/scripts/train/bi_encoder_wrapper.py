"""Drop-in for the hot-path half of the reference's ``scripts/train/bi_encoder_wrapper.py`` (kd-ensemble teachers).

* ``BiSparseModel.forward`` (reference :28-35): the fused sparse head without the L0 log, special-token columns zeroed.
* ``DenseModel`` (reference :38-59): CLS vector, L2-normalised (tiny; plain torch on the backbone output).
* ``BiEncoderWrapper.get_scores_batch`` (reference :117-146): per teacher q.d^T through the score kernel, row min-max
  normalisation + running mean + scale in one kernel per teacher (ops.minmax_accumulate).
* ``RemoteModel`` (DynamoDB-cached embeddings) is out of scope and raises.
"""
import logging

import torch

from ... import ops
from ..model.sparse_encoders import _split_mlm_backbone
from ..utils import gather_rep

logger = logging.getLogger(__name__)


class BiSparseModel(torch.nn.Module):
    @staticmethod
    def from_pretrained(path):
        return BiSparseModel(path)

    def __init__(self, model_id, backbone=None, tokenizer=None):
        super().__init__()
        import transformers
        self.backbone = backbone if backbone is not None else transformers.AutoModelForMaskedLM.from_pretrained(
            model_id, trust_remote_code=True)
        self.tokenizer = tokenizer if tokenizer is not None else transformers.AutoTokenizer.from_pretrained(model_id)
        self.special_token_ids = [self.tokenizer.vocab[t] for t in self.tokenizer.special_tokens_map.values()]
        self._split = None
        self._special_cols = {}

    def forward(self, **kwargs):
        if self._split is None:
            self._split = _split_mlm_backbone(self.backbone)
        hidden = self._split.transform(self._split.body(**kwargs)[0])
        dec = self._split.decoder
        values = ops.sparse_head(hidden, dec.weight, dec.bias, kwargs.get("attention_mask"), use_l0=False)
        key = (values.device.type, values.device.index)
        if key not in self._special_cols:
            self._special_cols[key] = torch.tensor(list(self.special_token_ids), dtype=torch.long, device=values.device)
        return values.index_fill(1, self._special_cols[key], 0.0)


class DenseModel(torch.nn.Module):
    @staticmethod
    def from_pretrained(path):
        return DenseModel(path)

    @staticmethod
    def get_dense_embedding(output):
        return torch.nn.functional.normalize(output[0][:, 0], p=2, dim=1)

    def __init__(self, model_id, backbone=None):
        super().__init__()
        import transformers
        self.backbone = backbone if backbone is not None else transformers.AutoModel.from_pretrained(
            model_id, trust_remote_code=True)

    def forward(self, **kwargs):
        return DenseModel.get_dense_embedding(self.backbone(**kwargs))


class RemoteModel(torch.nn.Module):
    @staticmethod
    def from_pretrained(path):
        return RemoteModel(path)

    def __init__(self, model_id):
        super().__init__()
        raise NotImplementedError("'remote' teachers (DynamoDB embedding cache) are outside the B200 hot path; "
                                  "use 'sparse' or 'dense' teachers")


class BiEncoderWrapper:
    CLS_MAP = {"sparse": BiSparseModel, "dense": DenseModel, "remote": RemoteModel}

    def __init__(self, types, model_ids, score_scale=30, use_in_batch_negatives=False, embedding_service=None,
                 models=None):
        assert len(types) == len(model_ids)
        assert len(types) != 0
        self.score_scale = score_scale
        self.use_in_batch_negatives = use_in_batch_negatives
        self.accelerator = None
        if models is None:
            models = [BiEncoderWrapper.CLS_MAP[t].from_pretrained(mid) for t, mid in zip(types, model_ids)]
        self.models = list(models)
        for m in self.models:
            m.eval()

    def get_scores_batch(self, q_features_list, d_features_list):
        assert len(q_features_list) == len(self.models)
        total = None
        share = float(self.score_scale) / len(self.models)
        with torch.no_grad():
            for model, qf, df in zip(self.models, q_features_list, d_features_list):
                q_rep = model(**qf).float()
                d_rep = model(**df).float()
                if self.use_in_batch_negatives:
                    d_rep = gather_rep(d_rep, self.accelerator)
                S = ops.scores(q_rep, d_rep, self.use_in_batch_negatives)
                total = ops.minmax_accumulate(S, total, scale=share)
        return total
