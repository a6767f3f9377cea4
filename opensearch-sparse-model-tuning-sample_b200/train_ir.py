"""Training entry point with the reference's command line (``train_ir.py config.yaml`` or ``--key value`` flags, also
under ``torchrun --nproc_per_node=N``) on the B200 hot path.

Same flow as the reference script (train_ir.py:30-150): parse the three argument groups, snapshot the config into
``output_dir``, build the sparse model, the loss functions from ``loss_types`` (weight ``ranking_loss_weight``,
``temperature``, ``use_in_batch_negatives``), AdamW with an optional separate learning rate for the IDF vector
(``idf_lr``), a linear warm-up/decay schedule, optional kd-ensemble teachers, then ``trainer.train()``.

Data (``data_type``): ``kd`` / ``posnegs`` read ``train_file`` through the ported dataset + collator classes
(``scripts/dataset``: a ``datasets.save_to_disk`` directory as upstream, or a .jsonl file with the same columns) and
need the model's tokenizer; ``synthetic`` / ``synthetic_posnegs`` generate seeded token batches (offline benchmarks);
``tensors`` loads a ``torch.save``d list of already collated batches. ``kd-ids`` (DynamoDB embedding service) raises.
"""
import logging
import os
import shutil
import sys

import torch
import torch.distributed as dist
import yaml

from .scripts import synthetic
from .scripts.args import parse_args
from .scripts.train.loss import LOSS_CLS_MAP
from .scripts.train.trainer import SparseModelTrainer
from .scripts.utils import get_model, set_logging

logger = logging.getLogger(__name__)


class _BatchList(torch.utils.data.Dataset):
    """Pre-collated batches; the DataLoader runs with batch_size=None semantics through an identity collator."""

    def __init__(self, batches):
        self.batches = batches

    def __len__(self):
        return len(self.batches)

    def __getitem__(self, i):
        return self.batches[i]


def build_dataset(data_args, training_args, model, rank):
    """-> (dataset, collate_fn, per-device batch size handed to the DataLoader)."""
    identity = lambda items: items[0]  # noqa: E731  (datasets of pre-collated batches)
    # posnegs rows carry 1 positive + sample_num_one_query negatives; kd rows carry sample_num_one_query docs
    G = data_args.sample_num_one_query + (1 if data_args.data_type == "synthetic_posnegs" else 0)
    if data_args.data_type in ("synthetic", "synthetic_posnegs"):
        needs_scores = any(t != "infonce" for t in data_args.loss_types) and not data_args.kd_ensemble_teacher_kwargs
        nq = training_args.per_device_train_batch_size
        n_scores = None
        if needs_scores:
            n_scores = G  # kd data carries the teacher scores of each query's own docs; in-batch scores come from teachers
        steps = max(1, training_args.max_steps)
        n_sets = 1 + len((data_args.kd_ensemble_teacher_kwargs or {}).get("types", []))
        batches = [synthetic.train_batch(nq, G, data_args.max_seq_length, query_len=32, seed=1234 + 1000 * rank + s,
                                         vocab_size=model.vocab_size, with_scores=n_scores, n_feature_sets=n_sets)
                   for s in range(min(steps, 64))]
        return _BatchList(batches), identity, 1
    if data_args.data_type == "tensors":
        return _BatchList(torch.load(data_args.train_file)), identity, 1
    from .scripts.dataset.collator import COLLATOR_CLS_MAP
    from .scripts.dataset.dataset import load_dataset
    if data_args.data_type not in COLLATOR_CLS_MAP:
        raise NotImplementedError(f"data_type={data_args.data_type!r}: supported are {sorted(COLLATOR_CLS_MAP)}, 'synthetic', "
                                  "'synthetic_posnegs' and 'tensors' ('kd-ids' needs the reference's DynamoDB service)")
    if data_args.train_file is None:
        raise ValueError("train_file must be specified (train_file_dir mixtures are not supported by this trainer)")
    if model.tokenizer is None or not callable(model.tokenizer):
        raise ValueError(f"data_type={data_args.data_type!r} tokenises text: the model needs a real tokenizer")
    dataset = load_dataset(path=data_args.train_file, cls=data_args.data_type, swap_times=data_args.swap_times,
                           sample_num_one_query=data_args.sample_num_one_query,
                           first_rank_thresh=data_args.first_rank_thresh)
    collator = COLLATOR_CLS_MAP[data_args.data_type](model.tokenizer, data_args.max_seq_length,
                                                     (data_args.kd_ensemble_teacher_kwargs or {}).get("teacher_tokenizer_ids", []))
    return dataset, collator, training_args.per_device_train_batch_size


def build_optimizer(model, model_args, data_args, training_args):
    """AdamW; the IDF vector gets its own learning rate when idf_lr is set and it is trainable; weight_decay applies to
    every group (reference train_ir.py:85-101). Fused implementation: it can skip a step on a device-side flag."""
    kw = dict(lr=training_args.learning_rate, weight_decay=training_args.weight_decay,
              betas=(training_args.adam_beta1, training_args.adam_beta2), eps=training_args.adam_epsilon,
              fused=next(model.parameters()).is_cuda)
    if not model_args.idf_requires_grad or data_args.idf_lr is None:
        opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], **kw)
    else:
        other = [p for p in model.parameters() if p is not model.idf_vector and p.requires_grad]
        opt = torch.optim.AdamW([{"params": [model.idf_vector], "lr": data_args.idf_lr}, {"params": other}], **kw)
        logger.info("idf_vector lr: %s", data_args.idf_lr)
    warm, total = training_args.warmup_steps, max(1, training_args.max_steps)

    def linear(step):
        if step < warm:
            return float(step) / float(max(1, warm))
        return max(0.0, float(total - step) / float(max(1, total - warm)))
    return opt, torch.optim.lr_scheduler.LambdaLR(opt, linear)


def main(argv=None, backbone=None, tokenizer=None, teacher_models=None):
    model_args, data_args, training_args = parse_args(argv)
    argv = sys.argv[1:] if argv is None else argv
    if int(training_args.gradient_accumulation_steps or 1) != 1:
        raise NotImplementedError("gradient_accumulation_steps != 1 is not supported by this trainer")
    if len(argv) == 1 and argv[0].endswith(".yaml"):
        shutil.copy(argv[0], os.path.join(training_args.output_dir, "train_config.yaml"))
    else:
        with open(os.path.join(training_args.output_dir, "config.yaml"), "w") as f:
            yaml.safe_dump({**vars(model_args), **vars(data_args), **vars(training_args)}, f)
    set_logging(training_args, "train.log")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    torch.manual_seed(training_args.seed)

    model = get_model(model_args, backbone=backbone, tokenizer=tokenizer).to(device)
    loss_functions = [LOSS_CLS_MAP[t](use_in_batch_negatives=data_args.use_in_batch_negatives,
                                      weight=data_args.ranking_loss_weight, temperature=data_args.temperature)
                      for t in data_args.loss_types]
    optimizers = build_optimizer(model, model_args, data_args, training_args)
    dataset, collator, loader_batch = build_dataset(data_args, training_args, model, int(os.environ.get("RANK", "0")))
    trainer = SparseModelTrainer(model_args, data_args, loss_functions, model=model, args=training_args,
                                 train_dataset=dataset, data_collator=collator, optimizers=optimizers)
    trainer.args.per_device_train_batch_size = loader_batch  # 1 when the dataset already yields collated batches
    if data_args.kd_ensemble_teacher_kwargs:
        trainer.set_bi_encoder_teacher(models=teacher_models)
    steps = trainer.train()
    logger.info("finished %d steps, ranking loss moving avg %s", steps, trainer.ranking_loss_moving_avg)
    return trainer


if __name__ == "__main__":
    main()
