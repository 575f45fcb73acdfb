"""First-contact check of the tcgen05 ToMe matching kernel against the FFMA path with identical bf16 rounding."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokenreduction_b200 import ops as T
torch.manual_seed(0)
for (b, n, r, d) in [(4, 197, 59, 64), (3, 138, 41, 64), (2, 97, 29, 64), (2, 50, 30, 32), (2, 197, 98, 64), (256, 197, 59, 64)]:
    m = torch.randn(b, n, d, device="cuda").bfloat16()
    a = T.tome_match(m, r, True, True, True)
    torch.cuda.synchronize()
    f = T.tome_match(m, r, True, True, False)
    torch.cuda.synchronize()
    same = [(x == y).float().mean().item() for x, y in zip(a, f)]
    print((b, n, r, d), "agreement unm/src/dst:", [round(v, 4) for v in same], flush=True)
