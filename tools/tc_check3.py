"""First-contact check of the 3xTF32 tensor-core distance path against the exact-fp32 FFMA path and float64."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokenreduction_b200 import ops as T
torch.manual_seed(0)
for (b, p, c) in [(2, 196, 384), (2, 49, 384), (3, 196, 768), (2, 60, 100), (2, 26, 64), (2, 130, 192), (150, 196, 384)]:
    x = torch.randn(b, p, c, device="cuda")
    d_tc = T.pairwise_dist(x, 1.0, False); torch.cuda.synchronize()
    d_ff = T.pairwise_dist(x, 1.0, True); torch.cuda.synchronize()
    d64 = torch.cdist(x.double(), x.double())
    d_at = torch.cdist(x, x)
    off = ~torch.eye(p, dtype=torch.bool, device="cuda")
    e = lambda d: float((d.double() - d64)[:, off].abs().max())
    print((b, p, c), "max |err| vs f64 offdiag: tc %.3e ffma %.3e aten %.3e | sym %s | diag max tc %.3e" %
          (e(d_tc), e(d_ff), e(d_at), bool(torch.equal(d_tc, d_tc.transpose(1, 2))), float(d_tc.diagonal(dim1=1, dim2=2).max())), flush=True)
