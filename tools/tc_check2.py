"""First-contact check of the tcgen05 soft-merge kernel against the FFMA path with identical bf16 rounding."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokenreduction_b200 import ops as T
torch.manual_seed(0)
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
for (b, p, k, c) in [(2, 196, 176, 768), (2, 176, 158, 768), (2, 158, 142, 384), (2, 64, 20, 128), (3, 60, 130, 200), (130, 196, 176, 768)]:
    x = torch.randn(b, p, c, device="cuda")
    v = torch.nn.functional.normalize(torch.randn(k, c, device="cuda"), dim=-1)
    o1, w1 = T.sinkhorn_merge(x, v, 1.0, 3, True, True); torch.cuda.synchronize()
    o2, w2 = T.sinkhorn_merge(x, v, 1.0, 3, True, False); torch.cuda.synchronize()
    print("sinkhorn", (b, p, k, c), "rel err out %.3e weights %.3e" % (rel(o1.float(), o2.float()), rel(w1, w2)), flush=True)
    lw, lb = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    q = torch.randn(k, c, device="cuda") * 0.05
    o1, w1 = T.patchmerger(x, lw, lb, q, 1.0, 1e-5, True, True); torch.cuda.synchronize()
    o2, w2 = T.patchmerger(x, lw, lb, q, 1.0, 1e-5, True, False); torch.cuda.synchronize()
    print("patchmerger", (b, p, k, c), "rel err out %.3e attn %.3e" % (rel(o1.float(), o2.float()), rel(w1, w2)), flush=True)
