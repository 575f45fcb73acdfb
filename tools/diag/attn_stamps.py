"""phase timeline of the fused attention kernel from the debug library's clock64 stamps (python -m tokenreduction_b200.build --stamps)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tokenreduction_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace(".so", "_dbg.so")
from tokenreduction_b200 import ops as T
lib = _lib.load()
B, N, H = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 197, 6
MODE = sys.argv[2] if len(sys.argv) > 2 else "plain"
kw = {"scores": dict(want_out=False, want_cls=True), "colsum": dict(want_cls=True, want_colsum=True), "plain": {}}[MODE]
qkv = torch.randn(B, N, 3 * H * 64, device="cuda").bfloat16()
for _ in range(3):
    T.attention(qkv, H, 0.125, **kw)
ncta = min(B * H, 2 * 148)
st = torch.zeros(ncta * 8 * 32, dtype=torch.int64, device="cuda")
lib.tokred_debug_set_stamps_attention.argtypes = [ctypes.c_void_p]
assert lib.tokred_debug_set_stamps_attention(st.data_ptr()) == 0
T.attention(qkv, H, 0.125, **kw)
torch.cuda.synchronize()
lib.tokred_debug_set_stamps_attention(None)
s = st.view(ncta, 8, 32).cpu().double()
names = ["start", "loaded", "S ready", "A done(up)", "B done(up)", "C done", "PV ready", "tile done", "A done(lo)", "B done(lo)"]
t0 = s[:, 0, 0:1]
for tile in range(2):
    rel = (s[:, tile, :10] - t0) / 1.9e3
    rel[s[:, tile, :10] == 0] = float("nan")
    print(f"tile {tile}: mean us since CTA start: " + " | ".join(f"{names[i]} {rel[:, i].nanmean().item():6.2f}" for i in range(10)))
    for cta in (0, 5, 150, 295):
        print(f"   CTA {cta}: " + " ".join(f"{rel[cta, i].item():6.2f}" for i in range(10)))
ends = (s[:, :8, 10] - t0) / 1.9e3
ends[s[:, :8, 10] == 0] = float("nan")
print("item end times (us since CTA start), mean over CTAs:", " ".join(f"{ends[:, i].nanmean().item():6.2f}" for i in range(8)))
for cta in (0, 5, 150, 295):
    print(f"   CTA {cta}: " + " ".join(f"{ends[cta, i].item():6.2f}" for i in range(8)))
