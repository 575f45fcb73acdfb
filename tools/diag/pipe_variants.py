"""each configuration in its own process (a CUDA fault poisons the context)."""
import subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CODE = r'''
import sys, torch
sys.path.insert(0, %r)
from tokenreduction_b200 import ops as T
kind, b, p, c = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
torch.manual_seed(0)
x = torch.randn(b, p, c, device="cuda")
for rep in range(3):
    if kind == "plain":
        d = T.pairwise_dist(x, 1.0, False)
    elif kind == "dpc":
        d = T.dpcknn_cluster(x, torch.rand(b, p, device="cuda"), max(p // 4, 1), 5, False)[0]
    else:
        d = T.kmedoids_fit(x, torch.rand(b, p, 1, device="cuda") + 5, max(p // 4, 1), 3, False)[0]
    torch.cuda.synchronize()
print("OK")
''' % ROOT
for cfg in [("plain", 3, 196, 384), ("plain", 300, 196, 384), ("dpc", 3, 196, 384), ("dpc", 256, 196, 384), ("kmed", 3, 196, 384),
            ("plain", 1, 49, 384), ("plain", 3, 128, 384), ("plain", 3, 129, 64), ("plain", 1, 196, 32), ("plain", 148, 196, 384)]:
    r = subprocess.run([sys.executable, "-c", CODE, *map(str, cfg)], capture_output=True, text=True, timeout=120)
    tail = (r.stdout + r.stderr).strip().splitlines()[-1][:150] if (r.stdout + r.stderr).strip() else ""
    print(cfg, "rc", r.returncode, tail, flush=True)
