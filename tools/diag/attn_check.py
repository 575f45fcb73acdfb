"""Bring-up check of the fused attention kernel against the eager bf16-autocast sequence on the same device."""
import sys, torch
sys.path.insert(0, ".")
from tokenreduction_b200 import ops as T


def eager(qkv, H, scale, bias=None, mask=None, ids=None):
    b, n, c3 = qkv.shape
    c = c3 // 3
    q, k, v = qkv.reshape(b, n, 3, H, c // H).permute(2, 0, 3, 1, 4)
    attn = (q @ k.transpose(-2, -1)) * scale
    if bias is not None:
        attn = attn + bias[:, None, None, :]
    if mask is not None:
        dm = mask.unsqueeze(1).unsqueeze(3) * mask.unsqueeze(1).unsqueeze(2)
        attn = attn.masked_fill(~dm, -torch.finfo(attn.dtype).max)
    attn = attn.float().softmax(-1)
    cls, colsum = attn[:, :, 0, :], None
    if ids is not None:
        attn = torch.gather(attn, 2, ids[:, None, :, None].expand(b, H, ids.shape[1], n))
    colsum = attn.sum(2)
    out = (attn.to(torch.bfloat16) @ v).transpose(1, 2).reshape(b, attn.shape[2], c)
    return out, cls, colsum


torch.manual_seed(0)
cases = [(2, 197, 6, 0), (3, 138, 6, 0), (2, 97, 12, 1), (2, 197, 6, 1), (2, 68, 6, 0), (1, 256, 3, 0), (2, 16, 2, 0), (2, 5, 1, 1),
         (3, 197, 6, 2), (3, 177, 12, 3), (2, 50, 6, 0), (2, 13, 6, 0)]
for (B, N, H, mode) in cases:
    qkv = (torch.randn(B, N, 3 * H * 64, device="cuda") * 1.5).bfloat16()
    kb = torch.rand(B, N, device="cuda").mul(3).add(1).log() if mode == 1 else None
    mask = ids = None
    if mode >= 2:
        mask = torch.rand(B, N, device="cuda") > 0.2
        mask[:, 0] = True
    if mode == 3:
        M = 143
        ids = torch.sort(torch.randint(0, N, (B, M), device="cuda"), dim=1).values
        ids[:, 0] = 0
        ids[:, -5:] = 0
    out, cls, cs = T.attention(qkv, H, 0.125, kb, mask, ids, True, True, True)
    ro, rc, rs = eager(qkv, H, 0.125, kb, mask, ids)
    torch.cuda.synchronize()
    eo = (out.float() - ro.float()).abs().max().item()
    ec = (cls - rc).abs().max().item()
    es = ((cs - rs).abs().max() / rs.abs().max()).item()
    _, cls2, _ = T.attention(qkv, H, 0.125, kb, mask, None, False, True, False)
    e2 = (cls2 - rc).abs().max().item()
    print(f"B={B} N={N} H={H} mode={mode}: out max|d|={eo:.3e} (ref max {ro.float().abs().max().item():.3f}) cls max|d|={ec:.3e} "
          f"scores-only cls {e2:.3e} colsum rel {es:.3e} nan={torch.isnan(out.float()).any().item()}")


def bench(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


import torch.nn.functional as F
for (B, N, H) in [(256, 197, 6), (256, 138, 6), (256, 97, 6), (256, 68, 6), (128, 197, 12), (1024, 197, 12)]:
    qkv = torch.randn(B, N, 3 * H * 64, device="cuda").bfloat16()
    q, k, v = qkv.reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    t0 = bench(lambda: T.attention(qkv, H, 0.125))
    t1 = bench(lambda: T.attention(qkv, H, 0.125, want_cls=True, want_colsum=True))
    t1a = bench(lambda: T.attention(qkv, H, 0.125, want_cls=True))
    t1b = bench(lambda: T.attention(qkv, H, 0.125, want_out=False, want_cls=True))
    t2 = bench(lambda: eager(qkv, H, 0.125)) if B * H * N * N * 4 < 8e9 else float("nan")
    t3 = bench(lambda: F.scaled_dot_product_attention(q, k, v))
    mb = B * N * 4 * H * 64 * 2 / 1e6
    print(f"B={B} N={N} H={H}: fused {t0:.1f} us ({mb / t0 * 1e-3 * 1e3:.0f} GB/s), +cls+colsum {t1:.1f} us, +cls {t1a:.1f} us, scores only {t1b:.1f} us, eager {t2:.1f} us, torch SDPA {t3:.1f} us")
