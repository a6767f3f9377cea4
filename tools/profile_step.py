"""Per-kernel device time of one eager training step (torch.profiler / CUPTI), padded vs padding-free body.
usage: python tools/profile_step.py [unpad_capacity ...]   (0 = padded)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torch.profiler import profile, ProfilerActivity


def run(capacity, top=90):
    wl = bench.WORKLOADS["c2"]
    bench.build_trainer.unpad_capacity = capacity if capacity > 0 else None
    bench.build_trainer.grad_sync = "ddp"
    dev = torch.device("cuda:0")
    trainer = bench.build_trainer(wl, "dense", dev)
    batches = [trainer._to_device(bench.host_batch(wl, 0, i), dev) for i in range(2)]
    clone = lambda b: {k: (list(v) if isinstance(v, list) else v) for k, v in b.items()}
    for i in range(4):
        trainer.training_step(clone(batches[i % 2]))
    torch.cuda.synchronize()
    n = 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(n):
            trainer.training_step(clone(batches[i % 2]))
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    total = sum(e.device_time_total for e in rows)
    print(f"== unpad_capacity={capacity}: {total / n / 1e3:.3f} ms of kernels per step, {sum(e.count for e in rows) / n:.0f} launches")
    for e in rows[:top]:
        print(f"{e.device_time_total / n / 1e3:8.3f} ms {e.count / n:6.0f}x  {e.key[:110]}")
    del trainer
    torch.cuda.empty_cache()


if __name__ == "__main__":
    for c in ([float(a) for a in sys.argv[1:]] or [0.0, 0.85]):
        run(c)
