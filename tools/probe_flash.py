import torch, time
try:
    import flash_attn
    from flash_attn import flash_attn_varlen_func
    print("flash_attn", flash_attn.__version__)
except Exception as e:
    print("flash_attn import failed:", repr(e)); raise SystemExit(0)
B, L, h, d = 160, 256, 12, 32
g = torch.Generator(device="cuda").manual_seed(0)
lens = torch.randint(L // 2, L + 1, (B,), device="cuda", generator=g)
cu = torch.zeros(B + 1, dtype=torch.int32, device="cuda"); cu[1:] = lens.cumsum(0)
T = int(cu[-1])
q = torch.randn(T, h, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
k = torch.randn(T, h, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
v = torch.randn(T, h, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
try:
    o = flash_attn_varlen_func(q, k, v, cu, cu, L, L, dropout_p=0.1, causal=False)
    o.sum().backward()
    torch.cuda.synchronize()
    print("varlen fwd+bwd ok", o.shape)
except Exception as e:
    print("varlen failed:", repr(e)); raise SystemExit(0)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
def fa():
    o = flash_attn_varlen_func(q, k, v, cu, cu, L, L, dropout_p=0.1, causal=False); o.backward(torch.ones_like(o))
print("flash varlen fwd+bwd ms", t(fa), "tokens", T, "of", B*L)
# padded SDPA with mask for comparison
qq = torch.randn(B, h, L, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
kk = torch.randn(B, h, L, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
vv = torch.randn(B, h, L, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
mask = (torch.arange(L, device="cuda")[None, :] < lens[:, None])
am = torch.zeros(B, 1, 1, L, device="cuda", dtype=torch.bfloat16).masked_fill(~mask[:, None, None, :], float("-inf"))
def sd():
    o = torch.nn.functional.scaled_dot_product_attention(qq, kk, vv, attn_mask=am, dropout_p=0.1); o.backward(torch.ones_like(o))
print("padded sdpa fwd+bwd ms", t(sd))
