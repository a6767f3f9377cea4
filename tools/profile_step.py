"""Per-kernel device time of eager training steps (torch.profiler / CUPTI). Single GPU or under torchrun (rank 0 prints).
usage: [torchrun --nproc-per-node N] tools/profile_step.py [--workload c2] [--unpad-capacity 0.85] [--grad-sync flat_overlap]
       [--rep-gather peer] [--regime dense] [--top 70]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--regime", default="dense")
    ap.add_argument("--unpad-capacity", type=float, default=0.85)
    ap.add_argument("--grad-sync", default="flat_overlap")
    ap.add_argument("--rep-gather", default="peer")
    ap.add_argument("--no-fused-body", action="store_true")
    ap.add_argument("--top", type=int, default=70)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = bench.WORKLOADS[args.workload]
    trainer = bench.build_trainer(wl, args, dev, grad_sync=args.grad_sync if world > 1 else "ddp",
                                  rep_gather=args.rep_gather if world > 1 else "nccl")
    teachers = bench.build_teachers(wl, dev)
    if teachers is not None:
        trainer.set_bi_encoder_teacher(models=teachers)
    batches = [trainer._to_device(bench.host_batch(wl, rank, i), dev) for i in range(2)]
    for i in range(4):
        trainer.training_step(bench.clone_inputs(batches[i % 2]))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        e0.record()
        for i in range(args.steps):
            trainer.training_step(bench.clone_inputs(batches[i % 2]))
        e1.record()
        torch.cuda.synchronize()
    if rank == 0:
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
        total = sum(e.device_time_total for e in rows)
        n = args.steps
        print(f"== {args.workload} world={world} grad_sync={trainer.grad_sync} rep_gather={trainer.rep_gather} "
              f"unpad_capacity={args.unpad_capacity}: {total / n / 1e3:.3f} ms of kernels per step (sum over streams), "
              f"{sum(e.count for e in rows) / n:.0f} launches, eager wall {e0.elapsed_time(e1) / n:.3f} ms/step")
        for e in rows[:args.top]:
            print(f"{e.device_time_total / n / 1e3:8.3f} ms {e.count / n:6.0f}x  {e.key[:110]}")
    if world > 1:
        dist.barrier()
        trainer.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
