"""one forward of a bench model for `ncu --metrics gpu__time_duration.sum` (launch list with device times)."""
import sys, torch
sys.path.insert(0, ".")
from argparse import Namespace
from tokenreduction_b200 import create_model
name, kr, b = (sys.argv[1], float(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("tome_small_patch16_224", 0.7, 256)
args = Namespace(keep_rate=[kr], reduction_loc=[3, 6, 9], distillation_type="none", k_neighbors=5, cluster_iters=3, sinkhorn_eps=1.0,
                 equal_weight=False, dyvit_distill=False)
torch.manual_seed(0)
m = create_model(name, num_classes=1000, args=args).eval().cuda()
x = torch.randn(b, 3, 224, 224, device="cuda")
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(3):
        m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
