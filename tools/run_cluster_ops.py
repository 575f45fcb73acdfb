"""Small driver for ncu captures of the DPC-KNN / K-Medoids kernels at the BASELINE stage-1 shape."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokenreduction_b200 import ops as T
b, p, k = 256, 196, 49
x = torch.randn(b, p, 384, device="cuda")
noise = torch.rand(b, p, device="cuda")
tw = torch.rand(b, p, 1, device="cuda") + 5.5
idx_token = torch.randint(0, p, (b, 196), device="cuda")
agg = torch.rand(b, 196, 1, device="cuda")
for _ in range(2):
    ic, _c = T.dpcknn_cluster(x, noise, k, 5)
    T.dpcknn_merge(x, idx_token, agg, ic, tw, k)
    T.kmedoids_fit(x, tw, k, 3, False)
torch.cuda.synchronize()
sizes = torch.stack([torch.bincount(ic[i], minlength=k) for i in range(b)])
print("cluster sizes: max per image mean %.1f, overall max %d" % (sizes.max(dim=1).values.float().mean().item(), sizes.max().item()))
