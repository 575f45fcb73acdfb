import sys, torch
sys.path.insert(0, ".")
from tokenreduction_b200 import ops as T
B, N, H = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 197, 6
qkv = torch.randn(B, N, 3 * H * 64, device="cuda").bfloat16()
for _ in range(4):
    T.attention(qkv, H, 0.125)
torch.cuda.synchronize()
