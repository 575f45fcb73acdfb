import sys, torch
sys.path.insert(0, "/root/repo")
from tokenreduction_b200 import ops as T
b, p, k, c = 128, 196, 176, 768
x = torch.randn(b, p, c, device="cuda")
v = torch.nn.functional.normalize(torch.randn(k, c, device="cuda"), dim=-1)
q = torch.randn(k, c, device="cuda") * 0.05
lw, lb = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
logits = torch.randn(b, p, k, device="cuda").bfloat16()
for _ in range(2):
    T.sinkhorn_merge(x, v, 1.0, 3, True, True)
    T.patchmerger(x, lw, lb, q, 1.0, 1e-5, True, True)
    T.sit_merge(x, logits, torch.ones(1, device="cuda"), True, True)
torch.cuda.synchronize()
