"""Phase breakdown (clock64 stamps of CTA 0) of the tensor-core soft-merge kernel."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokenreduction_b200 import _lib, ops as T
lib = _lib.load()
lib.tokred_debug_phase_buffer.argtypes = [ctypes.c_void_p]
buf = torch.zeros(8, dtype=torch.int64, device="cuda")
lib.tokred_debug_phase_buffer(buf.data_ptr())
b, p, k, c = 128, 196, 176, 768
x = torch.randn(b, p, c, device="cuda")
v = torch.nn.functional.normalize(torch.randn(k, c, device="cuda"), dim=-1)
q = torch.randn(k, c, device="cuda") * 0.05
lw, lb = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
logits = torch.randn(b, p, k, device="cuda").bfloat16()
names = ["stats", "gemm1", "Z epilogue", "W build (iters/softmax)", "gemm2+store"]
for label, fn in [("patchmerger", lambda: T.patchmerger(x, lw, lb, q, 1.0, 1e-5, True, True)),
                  ("sinkhorn", lambda: T.sinkhorn_merge(x, v, 1.0, 3, True, True)),
                  ("sit", lambda: T.sit_merge(x, logits, torch.ones(1, device="cuda"), True, True))]:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = buf.cpu().tolist()
    d = [s[i + 1] - s[i] for i in range(5)]
    tot = s[5] - s[0]
    print(label, "total cycles", tot, " | ".join(f"{n} {v_} ({100 * v_ / tot:.0f}%)" for n, v_ in zip(names, d)))
lib.tokred_debug_phase_buffer(None)
