"""diagnostic (test infrastructure: it uses the oracle, so it lives under tests/; not collected by pytest): where does the
tcgen05 bf16 matching differ from the oracle on decidable images?"""
import sys, os, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ops as O
import margins as MG
from tokenreduction_b200 import ops as T
g = lambda s: torch.Generator().manual_seed(s)
for n, r in ((197, 59), (138, 41), (97, 29)):
    b = 256
    metric = torch.randn(b, n, 64, generator=g(1400 + n)).bfloat16()
    unm_r, src_r, dst_r, nm_r = O.tome_match(metric, r, True, lowp=torch.bfloat16)
    for tc in (True, False):
        unm, src, dst = (t.cpu() for t in T.tome_match(metric.cuda(), r, True, True, tc))
        ok, sb = MG.tome_bf16_decidable(metric, r, True)
        same = (src == src_r).all(1) & (dst == dst_r).all(1) & (unm == unm_r).all(1)
        bad = (~same & ok).nonzero().flatten().tolist()
        print(f"N={n} tc={tc}: decidable {ok.float().mean():.3f} identical {same.float().mean():.3f} bad decidable images {bad}")
        m = metric.float(); m = m / m.norm(dim=-1, keepdim=True); mb = m.to(torch.bfloat16).double()
        s64 = mb[:, ::2] @ mb[:, 1::2].transpose(1, 2)
        # the oracle's own GPU/CPU fp32 matmul
        s32 = (mb[:, ::2].float() @ mb[:, 1::2].float().transpose(1, 2))
        for i in bad[:3]:
            dsrc = (src[i] != src_r[i]).nonzero().flatten().tolist()
            ddst = (dst[i] != dst_r[i]).nonzero().flatten().tolist()
            print("  image", i, "src diff at", dsrc[:6], "dst diff at", ddst[:6])
            for pos in (ddst or dsrc)[:4]:
                row_k, row_o = int(src[i, pos]), int(src_r[i, pos])
                ck, co = int(dst[i, pos]), int(dst_r[i, pos])
                print(f"    pos {pos}: kernel (row {row_k}, col {ck}) oracle (row {row_o}, col {co})")
                for (rr, cc) in ((row_k, ck), (row_o, co), (row_o, ck), (row_k, co)):
                    v = s64[i, rr, cc].item()
                    print(f"      s64[{rr},{cc}] = {v:.10f} bf16 {sb[i, rr, cc].item():.10f} fp32mm {s32[i, rr, cc].item():.10f} bdist {MG._bf16_boundary_distance(s64[i, rr, cc]).item():.3e}  rowmax_bf16 {sb[i, rr].max().item():.10f}")
