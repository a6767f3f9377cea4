"""Argument dataclasses with the reference's field names and defaults (``scripts/args.py``), so the same YAML files
(configs/config_infonce.yaml, config_kd.yaml, config_l0.yaml) parse unchanged.

``transformers.TrainingArguments`` needs ``accelerate`` (absent here), so ``TrainingArguments`` below is a plain
dataclass carrying the keys the configs use; unknown YAML keys are reported, not silently dropped.
"""
import dataclasses
import os
import sys
from dataclasses import dataclass, field
from typing import List, Optional, Union

import yaml

beir_datasets = "trec-covid,nfcorpus,nq,hotpotqa,fiqa,arguana,webis-touche2020,dbpedia-entity,scidocs,fever,climate-fever,scifact,quora"
miracl_datasets = "bn,te,es,fr,id,hi,ru,ar,zh,fa,ja,fi,sw,ko,en"


@dataclass
class DataTrainingArguments:
    max_seq_length: int = 512
    eval_max_seq_length: int = 512
    train_file: Optional[str] = None
    train_file_dir: Optional[str] = None
    data_type: Optional[str] = "kd"
    loss_types: List[str] = field(default_factory=lambda: ["kldiv"])
    beir_dir: str = "data/beir"
    miracl_dir: str = "mdata/miracl_eval"
    beir_datasets: str = beir_datasets
    miracl_datasets: str = miracl_datasets
    sample_num_one_query: int = 2
    use_in_batch_negatives: bool = False
    flops_d_lambda: float = 1e-3
    flops_d_T: float = 10000
    flops_q_lambda: Optional[float] = None
    flops_q_T: Optional[float] = None
    ranking_loss_weight: float = 1
    kd_ensemble_teacher_kwargs: Optional[Union[dict, str]] = field(default_factory=dict)
    idf_lr: Optional[float] = None
    first_rank_thresh: int = 10000
    use_two_phase: bool = False
    skip_ingest: bool = False
    do_search: bool = True
    query_prune: float = 0
    flops_threshold: Optional[int] = None
    swap_times: float = 0
    temperature: float = 1.0
    score_scale: float = 1.0


@dataclass
class ModelArguments:
    inf_free: bool = True
    model_name_or_path: Optional[str] = None
    tokenizer_name: Optional[str] = None
    idf_path: Optional[str] = None
    idf_requires_grad: Optional[bool] = False
    prune_ratio: Optional[float] = None
    preprocess_func: Optional[str] = None
    use_l0: bool = False
    # B200 extensions (absent from the reference's ModelArguments; defaults keep its behaviour)
    fuse_body: bool = True                    # backbone LayerNorm / Linear modules on the fused sm_100a kernels
    unpad_capacity: Optional[float] = None    # padding-free BERT body: rows = ceil(capacity * B * L); 1.0 never overflows
    attention: str = "auto"                   # attention kernels of the padding-free body: auto / own / flash

    def __post_init__(self):
        if self.tokenizer_name is None:
            self.tokenizer_name = self.model_name_or_path
        if self.idf_path == "null":
            self.idf_path = None
        if self.preprocess_func == "null":
            self.preprocess_func = None


@dataclass
class MiningArguments:
    mine_datasets: Optional[str] = None
    source: Optional[str] = None


@dataclass
class TrainingArguments:
    output_dir: str = "output/run"
    seed: int = 42
    learning_rate: float = 5e-5
    weight_decay: float = 0.0
    warmup_steps: int = 0
    max_steps: int = -1
    per_device_train_batch_size: int = 8
    per_device_eval_batch_size: int = 8
    logging_steps: int = 500
    log_level: str = "info"
    fp16: bool = False
    bf16: bool = False
    lr_scheduler_type: str = "linear"
    save_strategy: str = "steps"
    save_steps: int = 500
    dataloader_drop_last: bool = False
    dataloader_num_workers: int = 0
    max_grad_norm: Optional[float] = 1.0
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_epsilon: float = 1e-8
    gradient_accumulation_steps: int = 1
    save_safetensors: bool = True


def _coerce(field, value):
    """YAML 1.1 reads `1e-4` (no dot) as a string: numeric fields accept such strings, like HfArgumentParser does."""
    if isinstance(value, str):
        kind = str(field.type)
        try:
            if "float" in kind:
                return float(value)
            if "int" in kind and "Union" not in kind:
                return int(value)
        except ValueError:
            pass
    return value


def _split_config(cfg, classes):
    out, used = [], set()
    for cls in classes:
        fields = {f.name: f for f in dataclasses.fields(cls)}
        names = set(fields)
        out.append(cls(**{k: _coerce(fields[k], v) for k, v in cfg.items() if k in names}))
        used |= names & set(cfg)
    unknown = sorted(set(cfg) - used)
    return out, unknown


def parse_dict(cfg, classes=(ModelArguments, DataTrainingArguments, TrainingArguments)):
    parsed, unknown = _split_config(dict(cfg), classes)
    if unknown:
        import logging
        logging.getLogger(__name__).warning("ignoring config keys not used by this stack: %s", unknown)
    return tuple(parsed)


def parse_args(argv=None):
    """``script.py config.yaml`` or ``--key value`` flags -> (ModelArguments, DataTrainingArguments, TrainingArguments)."""
    argv = sys.argv[1:] if argv is None else list(argv)
    if len(argv) == 1 and argv[0].endswith(".yaml"):
        with open(os.path.abspath(argv[0])) as f:
            cfg = yaml.safe_load(f) or {}
    else:
        if len(argv) % 2 != 0:
            raise SystemExit("expected `config.yaml` or `--key value` pairs")
        cfg = {}
        for k, v in zip(argv[::2], argv[1::2]):
            if not k.startswith("--"):
                raise SystemExit(f"bad flag {k}")
            cfg[k[2:]] = yaml.safe_load(v)
    model_args, data_args, training_args = parse_dict(cfg)
    os.makedirs(training_args.output_dir, exist_ok=True)
    return model_args, data_args, training_args
