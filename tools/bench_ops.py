#!/usr/bin/env python
"""Per-op microbenchmark: every tokred kernel at the BASELINE.json stage shapes -> us, algorithmic GB/s, % of the
measured HBM roofline.  Runs on one B200:

    python tools/bench_ops.py [--out profiles/ops_rNN.json] [--iters 20] [--only tome]

Timing hygiene (B200_PROFILING.md): every iteration first READS a 1 GB buffer (> 126 MB L2: evicts the inputs and
keeps the GPU busy while the host enqueues; a read leaves the L2 full of CLEAN lines — an overwrite would leave
~100 MB of dirty lines whose write-back competes with the timed kernel), then records event / launches the kernel
through the C ABI / records event — so the bracket contains exactly one kernel and no host latency.
Median of the iterations after 3 warm-ups.  `--flush write` reproduces the pessimistic dirty-L2 variant.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import algorithmic_bytes, measured_peaks  # noqa: E402
from tokenreduction_b200 import _lib  # noqa: E402
from tokenreduction_b200 import ops as T  # noqa: E402

DEV = "cuda"


def g(seed):
    return torch.Generator().manual_seed(seed)


def spread_scores(b, p, seed):
    base = torch.linspace(0.05, 1.0, p)
    return torch.stack([base[torch.randperm(p, generator=g(seed + i))] for i in range(min(b, 16))]).repeat((b + 15) // 16, 1)[:b]


def cases():
    """(label, thunk returning nothing) — each thunk launches exactly one tokred kernel."""
    out = []

    def add(label, fn, *tensors):
        out.append((label, fn))

    # calibration: a 1-row gather = launch + event overhead of this timing method
    xs, ids1 = torch.randn(1, 4, 64, device=DEV), torch.zeros(1, 1, dtype=torch.int64, device=DEV)
    out.append(("calibration: empty launch (1-row gather)", lambda: T.gather_rows(xs, ids1)))
    # config 1: Top-K S kr 0.7 B=64 ; plus B=1024 large-batch sweep point
    for b in (64, 1024):
        for n, k in ((197, 137), (138, 96), (97, 67)):
            x = torch.randn(b, n, 384, device=DEV)
            s = spread_scores(b, n - 1, 1).to(DEV)
            out.append((f"topk_gather S B={b} N={n} k={k}", lambda x=x, s=s, k=k: T.topk_gather(x, s, k)))
    # config 2: ToMe S B=256 (bf16 metric, fp32 tokens) kr 0.7
    for n, r in ((197, 59), (138, 41), (97, 29)):
        b = 256
        m = torch.randn(b, n, 64, device=DEV).bfloat16()
        x = torch.randn(b, n, 384, device=DEV)
        size = torch.ones(b, n, 1, device=DEV)
        unm, src, dst = T.tome_match(m, r, True, True)
        out.append((f"tome_match S B={b} N={n} r={r} lowp tcgen05", lambda m=m, r=r: T.tome_match(m, r, True, True, True)))
        out.append((f"tome_match S B={b} N={n} r={r} lowp ffma", lambda m=m, r=r: T.tome_match(m, r, True, True, False)))
        out.append((f"tome_match S B={b} N={n} r={r} fp32", lambda m=m, r=r: T.tome_match(m.float(), r, True, False)))
        out.append((f"tome_merge S B={b} N={n} r={r}", lambda x=x, size=size, u=unm, s=src, d=dst: T.tome_merge(x, size, u, s, d, True, True)))
    # config 3: EViT / DynamicViT B kr 0.5, B=128 (8-GPU shard) and B=1024
    for b in (128, 1024):
        for n, k in ((197, 98), (100, 49), (51, 24)):
            x = torch.randn(b, n, 768, device=DEV)
            s = (spread_scores(b, n - 1, 2) / (n - 1)).to(DEV)
            out.append((f"evit_select_fuse B B={b} N={n} k={k}", lambda x=x, s=s, k=k: T.evit_select_fuse(x, s, k)))
        for n, k in ((197, 98), (99, 49), (50, 24)):
            x = torch.randn(b, n, 768, device=DEV)
            pred = torch.randn(b, n - 1, 2, device=DEV)
            out.append((f"dyvit keep (topk_gather) B B={b} N={n} k={k}", lambda x=x, p=pred, k=k: T.topk_gather(x, p[:, :, 0], k)))
            h = torch.randn(b, n - 1, 768, device=DEV).bfloat16()
            pol = torch.ones(b, n - 1, 1, device=DEV)
            out.append((f"dyvit_pool_concat B B={b} P={n - 1}", lambda h=h, pol=pol: T.dyvit_pool_concat(h, pol)))
    # config 4: DPC-KNN / K-Medoids S kr 0.25 B=256
    b = 256
    for p, k in ((196, 49), (49, 12), (12, 3)):
        x = torch.randn(b, p, 384, device=DEV)
        noise = torch.rand(b, p, device=DEV)
        tw = torch.rand(b, p, 1, device=DEV) + 5.5
        idx_token = torch.randint(0, p, (b, 196), device=DEV)
        agg = torch.rand(b, 196, 1, device=DEV)
        ic, _ = T.dpcknn_cluster(x, noise, k, 5)
        out.append((f"dpcknn_cluster S B={b} P={p} K={k} tf32x3 tcgen05", lambda x=x, nz=noise, k=k: T.dpcknn_cluster(x, nz, k, 5, False)))
        out.append((f"dpcknn_cluster S B={b} P={p} K={k} exact ffma", lambda x=x, nz=noise, k=k: T.dpcknn_cluster(x, nz, k, 5, True)))
        out.append((f"dpcknn_merge S B={b} P={p} K={k}", lambda x=x, it=idx_token, a=agg, ic=ic, tw=tw, k=k: T.dpcknn_merge(x, it, a, ic, tw, k)))
        out.append((f"kmedoids_fit S B={b} P={p} K={k} iters=3 tf32x3 tcgen05", lambda x=x, tw=tw, k=k: T.kmedoids_fit(x, tw, k, 3, False)))
        out.append((f"kmedoids_fit S B={b} P={p} K={k} iters=3 exact ffma", lambda x=x, tw=tw, k=k: T.kmedoids_fit(x, tw, k, 3, True)))
        attn = torch.softmax(torch.randn(b, 6, p + 1, p + 1, device=DEV), dim=-1)
        out.append((f"attn_colsum S B={b} N={p + 1}", lambda a=attn: T.attn_colsum(a, 1)))
    # config 5: Sinkhorn / PatchMerger / SiT / ATS B kr 0.9, B=128
    b = 128
    for p, k in ((196, 176), (176, 158), (158, 142)):
        x = torch.randn(b, p, 768, device=DEV)
        v = torch.nn.functional.normalize(torch.randn(k, 768, device=DEV), dim=-1)
        q = torch.randn(k, 768, device=DEV) * 0.05
        lw, lb = torch.ones(768, device=DEV), torch.zeros(768, device=DEV)
        logits = torch.randn(b, p, k, device=DEV).bfloat16()
        scale = torch.ones(1, device=DEV)
        for lowp, tc, tag in ((True, True, "lowp tcgen05"), (True, False, "lowp ffma"), (False, False, "fp32")):
            out.append((f"sinkhorn_merge B B={b} P={p} K={k} {tag}", lambda x=x, v=v, lowp=lowp, tc=tc: T.sinkhorn_merge(x, v, 1.0, 3, lowp, tc)))
            out.append((f"patchmerger B B={b} P={p} K={k} {tag}", lambda x=x, lw=lw, lb=lb, q=q, lowp=lowp, tc=tc: T.patchmerger(x, lw, lb, q, 1.0, 1e-5, lowp, tc)))
        out.append((f"sit_merge B B={b} P={p} K={k} lowp tcgen05", lambda x=x, l=logits, s=scale: T.sit_merge(x, l, s, True, True)))
        out.append((f"sit_merge B B={b} P={p} K={k} lowp ffma", lambda x=x, l=logits, s=scale: T.sit_merge(x, l, s, True, False)))
    # config 5 large-batch sweep (stage-1 shapes): several waves of per-image CTAs overlap their phases
    for b in (256, 512, 1024):
        p, k = 196, 176
        x = torch.randn(b, p, 768, device=DEV)
        v = torch.nn.functional.normalize(torch.randn(k, 768, device=DEV), dim=-1)
        q = torch.randn(k, 768, device=DEV) * 0.05
        lw, lb = torch.ones(768, device=DEV), torch.zeros(768, device=DEV)
        logits = torch.randn(b, p, k, device=DEV).bfloat16()
        scale = torch.ones(1, device=DEV)
        out.append((f"sweep sinkhorn_merge B B={b} P={p} K={k} lowp tcgen05", lambda x=x, v=v: T.sinkhorn_merge(x, v, 1.0, 3, True, True)))
        out.append((f"sweep patchmerger B B={b} P={p} K={k} lowp tcgen05", lambda x=x, lw=lw, lb=lb, q=q: T.patchmerger(x, lw, lb, q, 1.0, 1e-5, True, True)))
        out.append((f"sweep sit_merge B B={b} P={p} K={k} lowp tcgen05", lambda x=x, l=logits, s=scale: T.sit_merge(x, l, s, True, True)))
    # ATS: 8-GPU shard (B=128) and the single-GPU configuration of SURVEY section 8(d) (B=1024, stage 1)
    for b, shapes in ((128, ((197, 177), (177, 159), (159, 143))), (1024, ((197, 177),))):
        for n, count in shapes:
            attn = torch.softmax(4 * torch.randn(b, 12, n, n, device=DEV), dim=-1)
            v = torch.randn(b, 12, n, 64, device=DEV).bfloat16()
            mask = torch.ones(b, n, dtype=torch.bool, device=DEV)
            steps = torch.arange(1 / (2 * count), (2 * count - 1) / (2 * count), 2 / (2 * count)).to(DEV)   # models/ats.py:48
            ids, _, _ = T.ats_sample(v, attn, mask, steps)
            x = torch.randn(b, n, 768, device=DEV)
            tag = "" if b == 128 else "sweep "
            out.append((f"{tag}ats_sample B B={b} N={n} count={count}", lambda v=v, a=attn, m=mask, s=steps: T.ats_sample(v, a, m, s)))
            out.append((f"{tag}ats gather attn rows B B={b} N={n} M={count}", lambda a=attn, ids=ids: T.gather_rows(a, ids)))
            out.append((f"{tag}ats gather tokens B B={b} N={n} M={count}", lambda x=x, ids=ids: T.gather_rows(x, ids)))
            del attn, v, x
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "ops_r01.json"))
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--flush", default="read", choices=["read", "write"])
    a = ap.parse_args()
    peaks, kind = measured_peaks()
    flush = torch.zeros(256 << 20, dtype=torch.float32, device=DEV)      # 1 GB
    rows = []
    for label, fn in cases():
        if a.only and a.only not in label:
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        times, nbytes, kname = [], 0.0, ""
        for _ in range(a.iters):
            if a.flush == "read":
                flush.sum()
            else:
                flush.zero_()
            _lib.TIMELINE = []
            fn()
            tl, _lib.TIMELINE = _lib.TIMELINE, None
            torch.cuda.synchronize()
            assert len(tl) == 1, f"{label}: expected exactly one launch, got {len(tl)}"
            name, args, e0, e1 = tl[0]
            times.append(e0.elapsed_time(e1) * 1e3)
            nbytes, kname = algorithmic_bytes(name, args), name
        times.sort()
        med = times[len(times) // 2]
        gbs = nbytes / (med * 1e-6) / 1e9
        rows.append({"case": label, "kernel": kname.replace("tokred_", ""), "us_median": round(med, 2), "us_min": round(times[0], 2),
                     "alg_mb": round(nbytes / 1e6, 3), "alg_gbs": round(gbs, 1), "frac_hbm": round(gbs / peaks["hbm_gbs"], 4),
                     "roofline_us": round(nbytes / (peaks["hbm_gbs"] * 1e9) * 1e6, 2)})
        print(f"{label:60s} {med:9.2f} us  {nbytes / 1e6:9.2f} MB  {gbs:8.1f} GB/s  {100 * gbs / peaks['hbm_gbs']:5.1f}% of {kind} HBM peak", flush=True)
    with open(a.out, "w") as fh:
        json.dump({"peak_hbm_gbs": peaks["hbm_gbs"], "peak_source": kind, "gpu": torch.cuda.get_device_name(0),
                   "timing": f"CUDA events around one C-ABI launch, 1 GB L2 flush ({a.flush}) before each, median", "rows": rows}, fh, indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
