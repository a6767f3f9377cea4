"""Drop-in for the reference's ``scripts/model/sparse_encoders.py`` with the sparse head on sm_100a kernels.

Same public surface (SparseModel, SparseEncoder, SparsePostProcessor, sparse_embedding_to_query,
TokenizerWithProcessing, TextPreProcessors), same constructor arguments and attribute names, same outputs:
  * ``SparseModel._encode`` (reference :107-119): the backbone is split just before its vocabulary decoder; the
    decoder GEMM, mask multiply, max-pool over the sequence and log1p(relu) [+log1p] run as ONE fused tcgen05 kernel
    (ops.sparse_head), so the [B, L, V] logits are never materialised, and its backward is the sparse scatter kernel.
  * ``SparseModel._encode_inf_free`` (reference :121-127): ops.idf_query, bit-exact.
  * ``SparsePostProcessor`` (reference :130-150): GPU CSR compaction + one device->host copy.
There is no eager fallback: CPU tensors or an unsupported backbone head raise.
"""
import logging

import torch

from ... import ops

logger = logging.getLogger(__name__)


class TokenizerWithProcessing:
    """Applies a text pre-processing function before the wrapped tokenizer (reference :9-22)."""

    def __init__(self, original, process=None):
        self._original = original
        self.process = process

    def __call__(self, text, **kwargs):
        assert isinstance(text, list)
        assert isinstance(text[0], str)
        batch = text if self.process is None else self.process(text)
        return self._original(batch, **kwargs)

    def __getattr__(self, name):
        return getattr(self._original, name)


class TextPreProcessors:
    """Named text normalisers selectable through ``preprocess_func`` (reference :25-39)."""

    @staticmethod
    def to_lower(texts):
        return [t.lower() for t in texts]

    @staticmethod
    def blank_prefix(texts):
        return [" " + t for t in texts]

    @staticmethod
    def blank_prefix_lower(texts):
        return [" " + t.lower() for t in texts]


# ------------------------------------------------------------------------------------------- backbone splitting
class _HeadSplit:
    """How to run an MLM backbone up to (not including) its vocabulary decoder."""

    def __init__(self, body, transform, decoder):
        self.body, self.transform, self.decoder = body, transform, decoder


def _split_mlm_backbone(backbone):
    """Returns (body module, transform callable, decoder Linear) for the supported MLM head layouts."""
    if hasattr(backbone, "cls") and hasattr(backbone.cls, "predictions") and hasattr(backbone, "bert"):
        pred = backbone.cls.predictions  # BertForMaskedLM
        return _HeadSplit(backbone.bert, pred.transform, pred.decoder)
    if hasattr(backbone, "distilbert") and hasattr(backbone, "vocab_projector"):
        def transform(x, m=backbone):  # DistilBertForMaskedLM
            return m.vocab_layer_norm(m.activation(m.vocab_transform(x)))
        return _HeadSplit(backbone.distilbert, transform, backbone.vocab_projector)
    if hasattr(backbone, "lm_head") and hasattr(backbone.lm_head, "decoder") and hasattr(backbone.lm_head, "dense"):
        head = backbone.lm_head  # Roberta / XLM-R / Electra-style heads
        body = getattr(backbone, "roberta", None) or backbone.base_model

        def transform(x, h=head):
            x = h.dense(x)
            x = torch.nn.functional.gelu(x)
            return h.layer_norm(x)
        return _HeadSplit(body, transform, head.decoder)
    if hasattr(backbone, "sparse_b200_split"):
        return _HeadSplit(*backbone.sparse_b200_split())  # user-supplied (body, transform, decoder)
    raise NotImplementedError(
        f"{type(backbone).__name__}: unsupported masked-LM head layout. Supported: BertForMaskedLM (cls.predictions), "
        "DistilBertForMaskedLM (vocab_projector), Roberta-style lm_head, or a backbone exposing sparse_b200_split() -> "
        "(body, transform, decoder_linear). There is no unfused fallback.")


class _PruneFunction(torch.autograd.Function):
    """values * (values > prune_ratio * rowmax) with the mask treated as a constant (reference :115-119)."""

    @staticmethod
    def forward(ctx, rep, ratio):
        out = ops.prune_rows_(rep.detach().clone().contiguous(), ratio)
        ctx.save_for_backward(out != 0)
        return out

    @staticmethod
    def backward(ctx, g):
        (kept,) = ctx.saved_tensors
        return g * kept, None


class SparseModel(torch.nn.Module):
    def __init__(self, model_id, idf=None, tokenizer_id=None, idf_requires_grad=False, prune_ratio=None,
                 preprocess_func=None, use_l0=True, backbone=None, tokenizer=None, fuse_body=True, unpad_capacity=None,
                 attention="auto"):
        """Arguments as in the reference (:43-52). ``backbone``/``tokenizer`` may be passed pre-built (offline use:
        random-init architectures, tests, benchmarks); otherwise they are loaded with transformers as upstream.
        ``fuse_body`` swaps the backbone's LayerNorm / Linear modules for the fused sm_100a kernels (same parameters).
        ``unpad_capacity`` (None = off) runs a BERT body padding-free under autocast: real tokens are packed into
        ceil(capacity * B * L) rows (see packed_body.py; 1.0 never overflows, smaller values must cover the data);
        ``attention`` picks its attention kernels (auto / own / flash)."""
        super().__init__()
        import transformers
        if backbone is None:
            backbone = transformers.AutoModelForMaskedLM.from_pretrained(model_id, trust_remote_code=True)
        if tokenizer is None and (tokenizer_id or model_id) is not None:
            tokenizer = transformers.AutoTokenizer.from_pretrained(tokenizer_id or model_id)
        self.backbone = backbone
        self.fused_layers = 0
        if fuse_body:
            from .fused_layers import fuse_backbone
            self.fused_layers = fuse_backbone(self.backbone)
        self.__dict__["_packed"] = None  # kept out of the module tree: it only references the backbone's modules
        if unpad_capacity is not None:
            from .packed_body import PackedBertBody
            if PackedBertBody.supported(self.backbone):
                self.__dict__["_packed"] = PackedBertBody(self.backbone.bert, unpad_capacity, attention)
        self.tokenizer = tokenizer
        if preprocess_func is not None:
            func = getattr(TextPreProcessors, preprocess_func)
            logger.info("Using preprocess function %s: %s -> %s", preprocess_func, ["Hello WorldABC."], func(["Hello WorldABC."]))
            self.tokenizer = TokenizerWithProcessing(self.tokenizer, func)

        vocab = self.tokenizer.vocab
        self.special_token_ids = [vocab[tok] for tok in self.tokenizer.special_tokens_map.values()]
        self.vocab_size = len(vocab)
        try:  # gte-"new" backbones carry a larger embedding table than the tokenizer vocabulary (reference :73-84)
            rows = self.backbone.new.embeddings.word_embeddings.weight.shape[0]
            if rows != self.vocab_size:
                logger.info("reset the vocab size from %d to %d", self.vocab_size, rows)
                self.vocab_size = rows
        except AttributeError:
            pass

        weights = torch.ones(self.vocab_size, dtype=torch.float32)
        if idf is not None:
            logger.info("set idf to the model. requires_grad: %s", idf_requires_grad)
            to_id = self.tokenizer._convert_token_to_id_with_added_voc
            for token, value in idf.items():
                weights[to_id(token)] = value
        self.idf_vector = torch.nn.Parameter(weights, requires_grad=idf_requires_grad)
        self.idf_requires_grad = idf_requires_grad
        self.prune_ratio = prune_ratio
        self.use_l0 = use_l0
        self._split = None
        self._special_cache = {}
        logger.info("model prune ratio: %s, use l0: %s", self.prune_ratio, self.use_l0)

    # -- helpers -------------------------------------------------------------------------------------------------
    def _special_ids_on(self, device):
        key = (device.type, device.index)
        if key not in self._special_cache:
            self._special_cache[key] = torch.tensor(list(self.special_token_ids), dtype=torch.int32, device=device)
        return self._special_cache[key]

    def _use_packed(self, features):
        packed = self.__dict__.get("_packed")
        ids = features.get("input_ids")
        return (packed is not None and ids is not None and ids.is_cuda and torch.is_autocast_enabled("cuda")
                and torch.get_autocast_dtype("cuda") == torch.bfloat16)

    def head_inputs(self, **features):
        """Runs the backbone up to the decoder input: -> (hidden [B,L,H], decoder Linear)."""
        if self._split is None:
            self._split = _split_mlm_backbone(self.backbone)
        if self._use_packed(features):
            from .packed_body import PackedBertBody
            # [T_cap, H] bf16, real tokens only, MLM head transform included
            hidden, plan = self.__dict__["_packed"](head_transform=self.backbone.cls.predictions.transform, **features)
            return PackedBertBody.repad(hidden, plan), self._split.decoder
        seq = self._split.body(**features)[0]
        return self._split.transform(seq), self._split.decoder

    def unpad_overflows(self):
        """Number of batches whose real tokens did not fit the packed capacity so far (synchronises); 0 when off."""
        packed = self.__dict__.get("_packed")
        return 0 if packed is None or packed.overflow_count is None else int(packed.overflow_count)

    def unpad_counter(self):
        """The device-side overflow counter (int64 scalar tensor) or None."""
        packed = self.__dict__.get("_packed")
        return None if packed is None else packed.overflow_count

    def unpad_step_flag(self):
        """fp32 scalar (0-dim) device flag: 1.0 if a forward since the last unpad_step_reset() overflowed the packed capacity.
        None when the packed body is off or its capacity (>= 1.0) can never overflow."""
        packed = self.__dict__.get("_packed")
        return None if packed is None or packed.capacity >= 1.0 else packed.step_overflow

    def unpad_step_reset(self):
        packed = self.__dict__.get("_packed")
        if packed is not None and packed.step_overflow is not None:
            packed.step_overflow.zero_()

    # -- reference API -------------------------------------------------------------------------------------------
    def forward(self, inf_free=False, **kwargs):
        return self._encode_inf_free(**kwargs) if inf_free else self._encode(**kwargs)

    def _encode(self, _sink=None, **kwargs):
        """`_sink` (B200 extension, not a tokenizer feature): a scripts.peer.PeerSink -- the head kernel then stores its
        rows into every rank's gathered buffer as well (gather_rep fused into the GEMM epilogue)."""
        ids = kwargs.get("input_ids")
        if self._use_packed(kwargs) and ops.head_packed_supported(self.backbone.config.hidden_size, ids.shape[1]):
            # padding-free all the way: the head reads the packed [T, H] rows of the body through a 2-D tensor map and
            # its backward writes packed rows -- no padded [B, L, H] copy in either direction
            if self._split is None:
                self._split = _split_mlm_backbone(self.backbone)
            hidden, plan = self.__dict__["_packed"](head_transform=self.backbone.cls.predictions.transform, **kwargs)
            decoder = self._split.decoder
            rep = ops.sparse_head_packed(hidden, plan[4][:ids.shape[0] + 1], ids.shape[1], decoder.weight, decoder.bias,
                                         use_l0=self.use_l0, sink=_sink)
        else:
            hidden, decoder = self.head_inputs(**kwargs)
            rep = ops.sparse_head(hidden, decoder.weight, decoder.bias, kwargs.get("attention_mask"), use_l0=self.use_l0,
                                  sink=_sink)
        if self.prune_ratio is None:
            return rep
        return _PruneFunction.apply(rep, float(self.prune_ratio))

    def _encode_inf_free(self, _sink=None, **kwargs):
        input_ids = kwargs.get("input_ids")
        return ops.idf_query(input_ids, self.idf_vector, self._special_ids_on(input_ids.device))


class SparsePostProcessor(object):
    """Dense [B, V] rows -> list of {token: weight}; column 0 never appears (reference :130-150)."""

    def __init__(self, tokenizer):
        self.tokenizer = tokenizer
        self.id_to_token = ["" for _ in range(len(tokenizer.vocab) + 100)]
        for token, _id in tokenizer.vocab.items():
            self.id_to_token[_id] = token

    def __call__(self, sparse_vector):
        row_ptr, cols, vals = ops.compact_rows(sparse_vector, first_col=1)
        bounds = row_ptr.tolist()  # the single device->host synchronisation of the encode path
        total = bounds[-1]
        tokens = [self.id_to_token[i] for i in cols[:total].tolist()]
        weights = vals[:total].tolist()
        return [dict(zip(tokens[lo:hi], weights[lo:hi])) for lo, hi in zip(bounds[:-1], bounds[1:])]


class SparseEncoder:
    """Tokenise -> no-grad forward -> DF counting -> dict output (reference :153-181)."""

    def __init__(self, sparse_model, max_length, do_count=True):
        self.model = sparse_model
        self.tokenizer = sparse_model.tokenizer
        self.post_processor = SparsePostProcessor(tokenizer=sparse_model.tokenizer)
        self.do_count = do_count
        self.max_length = max_length
        self.device = next(self.model.backbone.parameters()).device
        self.reset_count()

    def reset_count(self):
        self._df = torch.zeros(self.model.vocab_size, dtype=torch.int64, device=self.device)

    @property
    def count_tensor(self):
        """Per-token document frequency as a float vector, like the reference attribute of the same name."""
        return self._df.to(torch.float32)

    def encode(self, texts, inf_free=False):
        features = self.tokenizer(list(texts), padding=True, truncation=True, return_tensors="pt",
                                  return_token_type_ids=False, max_length=self.max_length)
        features = {k: v.to(self.device) for k, v in features.items()}
        with torch.no_grad():
            output = self.model(inf_free=inf_free, **features)
        return self.encode_output(output)

    def encode_output(self, output):
        """Post-processing half of encode(): fused CSR compaction + DF count (reference :178-180)."""
        row_ptr, cols, vals = ops.compact_rows(output, first_col=1, df_count=self._df if self.do_count else None)
        bounds = row_ptr.tolist()
        total = bounds[-1]
        id_to_token = self.post_processor.id_to_token
        tokens = [id_to_token[i] for i in cols[:total].tolist()]
        weights = vals[:total].tolist()
        return [dict(zip(tokens[lo:hi], weights[lo:hi])) for lo, hi in zip(bounds[:-1], bounds[1:])]


def sparse_embedding_to_query(token_weight_map, field_name="text_sparse", query_prune=0):
    """OpenSearch neural_sparse query body with optional weight pruning (reference :184-194)."""
    if query_prune > 0:
        floor = max(token_weight_map.values()) * query_prune
        token_weight_map = {t: w for t, w in token_weight_map.items() if w > floor}
    return {"neural_sparse": {field_name: {"query_tokens": token_weight_map}}}
