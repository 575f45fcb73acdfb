"""Bring-up check of the fused attention kernel against the eager bf16-autocast sequence on the same device."""
import sys, torch
sys.path.insert(0, ".")
from tokenreduction_b200 import ops as T


def eager(qkv, H, scale, bias=None):
    b, n, c3 = qkv.shape
    c = c3 // 3
    q, k, v = qkv.reshape(b, n, 3, H, c // H).permute(2, 0, 3, 1, 4)
    attn = (q @ k.transpose(-2, -1)) * scale
    if bias is not None:
        attn = attn + bias[:, None, None, :]
    attn = attn.float().softmax(-1)
    out = (attn.to(torch.bfloat16) @ v).transpose(1, 2).reshape(b, n, c)
    return out, attn[:, :, 0, :]


torch.manual_seed(0)
for (B, N, H, bias) in [(2, 197, 6, False), (3, 138, 6, False), (2, 97, 12, True), (2, 197, 6, True), (2, 68, 6, False), (1, 256, 3, False), (2, 16, 2, False), (2, 5, 1, True)]:
    qkv = (torch.randn(B, N, 3 * H * 64, device="cuda") * 1.5).bfloat16()
    kb = torch.rand(B, N, device="cuda").mul(3).add(1).log() if bias else None
    out, cls = T.attention(qkv, H, 0.125, kb, True)
    ro, rc = eager(qkv, H, 0.125, kb)
    torch.cuda.synchronize()
    eo = (out.float() - ro.float()).abs().max().item()
    ec = (cls - rc).abs().max().item()
    print(f"B={B} N={N} H={H} bias={bias}: out max|d|={eo:.3e} (ref max {ro.float().abs().max().item():.3f}) cls max|d|={ec:.3e} "
          f"nan={torch.isnan(out.float()).any().item()}")
# timing
import time
B, N, H = 256, 197, 6
qkv = torch.randn(B, N, 3 * H * 64, device="cuda").bfloat16()
for _ in range(3):
    T.attention(qkv, H, 0.125, None, False)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    T.attention(qkv, H, 0.125, None, False)
e1.record(); torch.cuda.synchronize()
print(f"fused B={B} N={N} H={H}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
with torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(3):
        eager(qkv, H, 0.125)
    e0.record()
    for _ in range(20):
        eager(qkv, H, 0.125)
    e1.record(); torch.cuda.synchronize()
print(f"eager: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
import torch.nn.functional as F
q, k, v = qkv.reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
for _ in range(3):
    F.scaled_dot_product_attention(q, k, v)
e0.record()
for _ in range(20):
    F.scaled_dot_product_attention(q, k, v)
e1.record(); torch.cuda.synchronize()
print(f"torch SDPA: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
