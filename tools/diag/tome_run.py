import sys, torch
sys.path.insert(0, ".")
from tokenreduction_b200 import ops as T
B, N, H = 256, 197, 6
qkv = torch.randn(B, N, 3 * H * 64, device="cuda").bfloat16()
for _ in range(4):
    u, s, d = T.tome_match_qkv(qkv, H, 59, True)
torch.cuda.synchronize()
