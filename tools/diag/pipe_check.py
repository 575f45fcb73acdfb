"""diagnostic for the pipelined distance kernels: shapes one by one, synchronised, against torch.cdist."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tokenreduction_b200 import ops as T
torch.manual_seed(0)
shapes = [(3, 196, 384), (1, 49, 384), (2, 130, 192), (5, 208, 96), (300, 196, 384), (3, 60, 100), (4, 26, 64)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
for b, p, c in shapes:
    x = torch.randn(b, p, c, device="cuda")
    d = T.pairwise_dist(x, 1.0, False)
    torch.cuda.synchronize()
    d64 = torch.cdist(x.double(), x.double())
    off = ~torch.eye(p, dtype=torch.bool, device="cuda")
    err = (d.double() - d64)[:, off].abs().max().item()
    ref = (torch.cdist(x, x).double() - d64)[:, off].abs().max().item()
    print(f"B={b} P={p} C={c}: max err {err:.3e} (ATen fp32 cdist {ref:.3e}) symmetric {torch.equal(d, d.transpose(1, 2))} "
          f"per-image worst {[(round(v, 8)) for v in (d.double() - d64).abs().flatten(1).max(1).values[:6].tolist()]}", flush=True)
