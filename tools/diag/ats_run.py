"""one ats_sample launch at the BASELINE stage-1 shape (for ncu)."""
import sys, torch
sys.path.insert(0, ".")
from tokenreduction_b200 import ops as T
b, h, n, dh = 128, 12, 197, 64
torch.manual_seed(0)
v = torch.randn(b, h, n, dh, device="cuda").bfloat16()
attn = torch.softmax(torch.randn(b, h, n, device="cuda"), -1)
mask = torch.ones(b, n, dtype=torch.bool, device="cuda")
count = 177
steps = torch.arange(1 / (2 * count), (2 * count - 1) / (2 * count), 2 / (2 * count)).cuda()      # models/ats.py:48
for _ in range(3):
    out = T.ats_sample(v, attn, mask, steps)
torch.cuda.synchronize()
