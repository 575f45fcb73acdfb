"""phase timeline of the pipelined DPC-KNN kernel from the debug library's clock64 stamps (python -m tokenreduction_b200.build --stamps)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tokenreduction_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace(".so", "_dbg.so")
from tokenreduction_b200 import ops as T
lib = _lib.load()
b, p, c, k = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 196, 384, 49
x = torch.randn(b, p, c, device="cuda"); noise = torch.rand(b, p, device="cuda")
for _ in range(3):
    T.dpcknn_cluster(x, noise, k, 5)
ncta = min(b, 148)
st = torch.zeros(ncta * 8 * 32, dtype=torch.int64, device="cuda")
lib.tokred_debug_set_stamps_cluster.argtypes = [ctypes.c_void_p]
assert lib.tokred_debug_set_stamps_cluster(st.data_ptr()) == 0
T.dpcknn_cluster(x, noise, k, 5)
torch.cuda.synchronize()
lib.tokred_debug_set_stamps_cluster(None)
s = st.view(ncta, 8, 32).cpu()
names = {0: "F first stage start", 1: "F first MMA issued", 2: "F last MMA issued", 8: "B wait acc_full", 9: "B acc_full seen", 10: "B drain done",
         11: "B knn+density done", 12: "B delta/score done", 13: "B rank done", 14: "B image done"}
for cta in (0, 1, 120, 147):
    if cta >= ncta: continue
    t0 = int(s[cta, 0, 0])
    print(f"CTA {cta}:")
    for img in range(2):
        if int(s[cta, img, 0]) == 0: continue
        print("  image", img, " | ".join(f"{names[i]} {(int(s[cta, img, i]) - t0) / 1.9e3:7.2f}us" for i in sorted(names) if int(s[cta, img, i])))
    print("  loader thread 0 totals over all chunks: wait stage_free %.2fus, wait loads %.2fus, norms+convert+STS+arrive %.2fus" % tuple(int(s[cta, 7, i]) / 1.9e3 for i in range(3)))
    if cta == 120:
        for c in range(12):
            print("   chunk %2d: loader0 stage_free seen %6.2f  loader0 staged+arrived %6.2f | MMA warp stage_full seen %6.2f  MMAs issued+committed %6.2f" % (
                c, *[(int(s[cta, k, c]) - t0) / 1.9e3 for k in (4, 5, 2, 3)]))
