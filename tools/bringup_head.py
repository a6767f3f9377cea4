"""GPU bring-up of the fused head: compares against an fp32 torch matmul on the same bf16-rounded inputs."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_b200
from sparse_b200 import ops

def ref(hidden, W, bias, mask, l0):
    logits = hidden.float() @ W.float().t()
    if bias is not None:
        logits = logits + bias
    vals, idx = torch.max(logits * mask.unsqueeze(-1), dim=1)
    rep = torch.log1p(torch.relu(vals))
    if l0:
        rep = torch.log1p(rep)
    return rep, vals, idx

def case(B, L, H, V, ragged=True, l0=False, bias_shift=0.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    hidden = torch.randn(B, L, H, device="cuda", generator=g).bfloat16()
    W = (torch.randn(V, H, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(V, device="cuda", generator=g) * 0.1 + bias_shift
    if ragged:
        lens = torch.randint(max(1, L // 2), L + 1, (B,), device="cuda", generator=g)
    else:
        lens = torch.full((B,), L, device="cuda")
    mask = (torch.arange(L, device="cuda")[None, :] < lens[:, None]).long()
    rep, xmax, amax = ops.head_forward(hidden, W, bias, mask, use_l0=l0)
    torch.cuda.synchronize()
    r_rep, r_vals, r_idx = ref(hidden, W, bias, mask, l0)
    err = (rep - r_rep).abs().max().item()
    errx = (xmax - r_vals).abs().max().item()
    act = r_rep > 0
    agree = (amax.long() == r_idx)[act].float().mean().item() if act.any() else 1.0
    agree_all = (amax.long() == r_idx).float().mean().item()
    print(f"B={B} L={L} H={H} V={V} ragged={ragged} l0={l0}: max|rep err|={err:.3e} max|xmax err|={errx:.3e} "
          f"argmax agree(active)={agree:.6f} (all)={agree_all:.6f} active={act.float().mean().item():.3f}", flush=True)
    return err

def bench(B, L, H, V, iters=20, ragged=False, reps=5):
    g = torch.Generator(device="cuda").manual_seed(1)
    hidden = torch.randn(B, L, H, device="cuda", generator=g).bfloat16()
    W = (torch.randn(V, H, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.zeros(V, device="cuda")
    if ragged:
        lens = torch.randint(L // 2, L + 1, (B,), device="cuda", generator=g)
    else:
        lens = torch.full((B,), L, device="cuda")
    mask = (torch.arange(L, device="cuda")[None, :] < lens[:, None]).long()
    for _ in range(5):
        ops.head_forward(hidden, W, bias, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for _ in range(reps):
        e0.record()
        for _ in range(iters):
            ops.head_forward(hidden, W, bias, mask)
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / iters)
    ms = min(times)
    fl = 2.0 * B * L * H * V
    print(f"bench B={B} L={L} H={H} ragged={ragged}: min {ms:.3f} ms median {sorted(times)[len(times)//2]:.3f} ms  "
          f"{fl/ms/1e9:.1f} TFLOP/s (all positions counted)", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    if len(sys.argv) > 1 and sys.argv[1] == "bench":
        bench(160, 256, 384, 30522)
        bench(160, 256, 384, 30522, ragged=True)
        bench(64, 512, 768, 30522)
        bench(64, 512, 768, 30522, ragged=True)
        bench(256, 128, 384, 30522)
        sys.exit(0)
    case(2, 128, 64, 128, ragged=False)
    case(2, 128, 384, 1000, ragged=False)
    case(8, 128, 384, 30522)
    case(5, 100, 384, 30522)
    case(3, 256, 768, 30522, l0=True)
    case(3, 512, 384, 5000, l0=True)
    case(2, 300, 384, 5000)
    case(6, 512, 384, 30522, l0=True)
    case(3, 1000, 128, 30522)
    case(4, 37, 128, 3000, bias_shift=-1.0)
    if len(sys.argv) > 1 and sys.argv[1] == "bench":
        pass
    bench(160, 256, 384, 30522)
    bench(160, 256, 384, 30522, ragged=True)
    bench(64, 512, 768, 30522)
    bench(64, 512, 768, 30522, ragged=True)
    bench(256, 128, 384, 30522)
