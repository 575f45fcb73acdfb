#!/usr/bin/env python
"""Per-op microbenchmark: every tokred kernel at the BASELINE.json stage shapes -> us, algorithmic GB/s, % of the
measured HBM roofline.  Runs on one B200:

    python tools/bench_ops.py [--out profiles/ops_rNN.json] [--iters 20] [--only tome]

Timing hygiene (B200_PROFILING.md; VERDICT r1 weak #10: a one-launch event bracket has an ~8 us floor and ~2 us
quantisation): N launches per CUDA-event pair, enqueued behind a device-side sleep so that no host latency sits inside
the bracket; the launches cycle over replicas of the inputs whose total size exceeds twice the 126 MB L2 (inputs are
never L2-resident; cases too small for that are flagged), and the outputs of the last round stay referenced so that
every launch writes fresh memory.  Per-launch time = elapsed / N, median over the event pairs.
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
    """(label, make) — make() allocates a fresh set of input tensors and returns a thunk that launches exactly one
    tokred kernel on them (several replicas of every case are cycled so that consecutive launches never find their
    inputs in the 126 MB L2)."""
    out = []

    def add(label, make):
        out.append((label, make))

    def rnd(*shape, dtype=torch.float32):
        return torch.randn(*shape, device=DEV).to(dtype)

    # config 1: Top-K S kr 0.7 B=64 ; plus B=1024 large-batch sweep point
    for b in (64, 1024):
        for n, k in ((197, 137), (138, 96), (97, 67)):
            def mk(b=b, n=n, k=k):
                x, s = rnd(b, n, 384), spread_scores(b, n - 1, 1).to(DEV)
                return lambda: T.topk_gather(x, s, k)
            add(f"topk_gather S B={b} N={n} k={k}", mk)
    # config 2: ToMe S B=256 (bf16 metric, fp32 tokens) kr 0.7
    for n, r in ((197, 59), (138, 41), (97, 29)):
        b = 256

        def mk_match(mode, b=b, n=n, r=r):
            m = rnd(b, n, 64, dtype=torch.bfloat16)
            if mode == "fp32":
                m = m.float()
                return lambda: T.tome_match(m, r, True, False)
            return lambda: T.tome_match(m, r, True, True, mode == "tc")

        def mk_merge(b=b, n=n, r=r):
            m, x, size = rnd(b, n, 64, dtype=torch.bfloat16), rnd(b, n, 384), torch.ones(b, n, 1, device=DEV)
            unm, src, dst = T.tome_match(m, r, True, True)
            return lambda: T.tome_merge(x, size, unm, src, dst, True, True)
        def mk_merge_ln(b=b, n=n, r=r):
            m, x, size = rnd(b, n, 64, dtype=torch.bfloat16), rnd(b, n, 384), torch.ones(b, n, 1, device=DEV)
            br, gam, bet = rnd(b, n, 384, dtype=torch.bfloat16), rnd(384), rnd(384)
            unm, src, dst = T.tome_match(m, r, True, True)
            return lambda: T.tome_merge_ln(x, br, size, unm, src, dst, gam, bet, 1e-6, True)

        def mk_match_qkv(b=b, n=n, r=r):
            qkv = rnd(b, n, 3 * 384, dtype=torch.bfloat16)
            return lambda: T.tome_match_qkv(qkv, 6, r, True)
        add(f"tome_match S B={b} N={n} r={r} from qkv keys (head mean in-kernel) tcgen05", mk_match_qkv)
        add(f"tome_match S B={b} N={n} r={r} lowp tcgen05", lambda f=mk_match: f("tc"))
        add(f"tome_match S B={b} N={n} r={r} lowp ffma", lambda f=mk_match: f("ffma"))
        add(f"tome_match S B={b} N={n} r={r} fp32", lambda f=mk_match: f("fp32"))
        add(f"tome_merge S B={b} N={n} r={r}", mk_merge)
        add(f"tome_merge_ln S B={b} N={n} r={r} (x + attn branch, merge, norm2 in one launch)", mk_merge_ln)
    # config 3: EViT / DynamicViT B kr 0.5, B=128 (8-GPU shard) and B=1024
    for b in (128, 1024):
        for n, k in ((197, 98), (100, 49), (51, 24)):
            def mk(b=b, n=n, k=k):
                x, s = rnd(b, n, 768), (spread_scores(b, n - 1, 2) / (n - 1)).to(DEV)
                return lambda: T.evit_select_fuse(x, s, k)
            add(f"evit_select_fuse B B={b} N={n} k={k}", mk)
        for n, k in ((197, 98), (99, 49), (50, 24)):
            def mk_keep(b=b, n=n, k=k):
                x, pred = rnd(b, n, 768), rnd(b, n - 1, 2)
                return lambda: T.topk_gather(x, pred[:, :, 0], k)

            def mk_pool(b=b, n=n):
                h, pol = rnd(b, n - 1, 768, dtype=torch.bfloat16), torch.ones(b, n - 1, 1, device=DEV)
                return lambda: T.dyvit_pool_concat(h, pol)
            add(f"dyvit keep (topk_gather) B B={b} N={n} k={k}", mk_keep)
            add(f"dyvit_pool_concat B B={b} P={n - 1}", mk_pool)
    # config 4: DPC-KNN / K-Medoids S kr 0.25 B=256
    b = 256
    for p, k in ((196, 49), (49, 12), (12, 3)):
        def mk_dpc(exact, b=b, p=p, k=k):
            x, noise = rnd(b, p, 384), torch.rand(b, p, device=DEV)
            return lambda: T.dpcknn_cluster(x, noise, k, 5, exact)

        def mk_dmerge(b=b, p=p, k=k):
            x, noise, tw = rnd(b, p, 384), torch.rand(b, p, device=DEV), torch.rand(b, p, 1, device=DEV) + 5.5
            idx_token, agg = torch.randint(0, p, (b, 196), device=DEV), torch.rand(b, 196, 1, device=DEV)
            ic, _ = T.dpcknn_cluster(x, noise, k, 5)
            return lambda: T.dpcknn_merge(x, idx_token, agg, ic, tw, k)

        def mk_kmed(exact, b=b, p=p, k=k):
            x, tw = rnd(b, p, 384), torch.rand(b, p, 1, device=DEV) + 5.5
            return lambda: T.kmedoids_fit(x, tw, k, 3, exact)

        def mk_colsum(b=b, p=p):
            attn = torch.softmax(rnd(b, 6, p + 1, p + 1), dim=-1)
            return lambda: T.attn_colsum(attn, 1)
        add(f"dpcknn_cluster S B={b} P={p} K={k} tf32x3 tcgen05", lambda f=mk_dpc: f(False))
        add(f"dpcknn_cluster S B={b} P={p} K={k} exact ffma", lambda f=mk_dpc: f(True))
        add(f"dpcknn_merge S B={b} P={p} K={k}", mk_dmerge)
        add(f"kmedoids_fit S B={b} P={p} K={k} iters=3 tf32x3 tcgen05", lambda f=mk_kmed: f(False))
        add(f"kmedoids_fit S B={b} P={p} K={k} iters=3 exact ffma", lambda f=mk_kmed: f(True))
        add(f"attn_colsum S B={b} N={p + 1}", mk_colsum)
    # config 5: Sinkhorn / PatchMerger / SiT B kr 0.9: 8-GPU shard (B=128, all stages) and the sweep {256, 512, 1024} (stage 1)
    for b, shapes, tags in ((128, ((196, 176), (176, 158), (158, 142)), (("lowp tcgen05", True, True), ("lowp ffma", True, False), ("fp32", False, False))),
                            (256, ((196, 176),), (("lowp tcgen05", True, True),)), (512, ((196, 176),), (("lowp tcgen05", True, True),)),
                            (1024, ((196, 176),), (("lowp tcgen05", True, True),))):
        pre = "" if b == 128 else "sweep "
        for p, k in shapes:
            for tag, lowp, tc in tags:
                def mk_sink(b=b, p=p, k=k, lowp=lowp, tc=tc):
                    x, v = rnd(b, p, 768), torch.nn.functional.normalize(rnd(k, 768), dim=-1)
                    return lambda: T.sinkhorn_merge(x, v, 1.0, 3, lowp, tc)

                def mk_pm(b=b, p=p, k=k, lowp=lowp, tc=tc):
                    x, q = rnd(b, p, 768), rnd(k, 768) * 0.05
                    lw, lb = torch.ones(768, device=DEV), torch.zeros(768, device=DEV)
                    return lambda: T.patchmerger(x, lw, lb, q, 1.0, 1e-5, lowp, tc)
                add(f"{pre}sinkhorn_merge B B={b} P={p} K={k} {tag}", mk_sink)
                add(f"{pre}patchmerger B B={b} P={p} K={k} {tag}", mk_pm)
                if lowp:
                    def mk_sit(b=b, p=p, k=k, tc=tc):
                        x, logits, scale = rnd(b, p, 768), rnd(b, p, k, dtype=torch.bfloat16), torch.ones(1, device=DEV)
                        return lambda: T.sit_merge(x, logits, scale, True, tc)
                    add(f"{pre}sit_merge B B={b} P={p} K={k} {tag}", mk_sit)
    # ATS: 8-GPU shard (B=128) and the single-GPU configuration of SURVEY section 8(d) (B=1024, stage 1)
    for b, shapes in ((128, ((197, 177), (177, 159), (159, 143))), (1024, ((197, 177),))):
        for n, count in shapes:
            tag = "" if b == 128 else "sweep "

            def mk_ats(kind, b=b, n=n, count=count):
                attn = torch.softmax(4 * rnd(b, 12, n, n), dim=-1)
                v = rnd(b, 12, n, 64, dtype=torch.bfloat16)
                mask = torch.ones(b, n, dtype=torch.bool, device=DEV)
                steps = torch.arange(1 / (2 * count), (2 * count - 1) / (2 * count), 2 / (2 * count)).to(DEV)   # models/ats.py:48
                if kind == "sample":
                    return lambda: T.ats_sample(v, attn, mask, steps)
                ids, _, _ = T.ats_sample(v, attn, mask, steps)
                if kind == "rows":
                    del v
                    return lambda: T.gather_rows(attn, ids)
                del attn, v
                x = rnd(b, n, 768)
                return lambda: T.gather_rows(x, ids)
            add(f"{tag}ats_sample B B={b} N={n} count={count}", lambda f=mk_ats: f("sample"))
            add(f"{tag}ats gather attn rows B B={b} N={n} M={count}", lambda f=mk_ats: f("rows"))
            add(f"{tag}ats gather tokens B B={b} N={n} M={count}", lambda f=mk_ats: f("tokens"))
    # f1: attention producer (no [B,H,N,N] tensor) at the stage sizes of configs 2 (DeiT-S, B=256) and 3/5 (DeiT-B)
    for b, h, sizes in ((256, 6, (197, 138, 97, 68)), (128, 12, (197,)), (1024, 12, (197, 100, 51))):
        for n in sizes:
            def mk_attn(b=b, h=h, n=n, side=False):
                qkv = rnd(b, n, 3 * h * 64, dtype=torch.bfloat16)
                return lambda: T.attention(qkv, h, 0.125, None, None, None, True, side, side)
            tag = "S" if h == 6 else "B"
            add(f"attention {tag} B={b} N={n} H={h}", mk_attn)
            if n == 197 and b != 1024:
                add(f"attention {tag} B={b} N={n} H={h} + cls rows + column sums", lambda f=mk_attn: f(side=True))

    def mk_attn_bias(b=256, h=6, n=197):
        qkv, bias = rnd(b, n, 3 * h * 64, dtype=torch.bfloat16), torch.rand(b, n, device=DEV)
        return lambda: T.attention(qkv, h, 0.125, bias)
    add("attention S B=256 N=197 H=6 + ToMe log-size bias", mk_attn_bias)

    def mk_attn_ats(b=128, h=12, n=197, m=177):
        qkv = rnd(b, n, 3 * h * 64, dtype=torch.bfloat16)
        mask = torch.ones(b, n, dtype=torch.bool, device=DEV)
        ids = torch.arange(m, device=DEV).expand(b, -1).contiguous()
        return lambda: T.attention(qkv, h, 0.125, None, mask, ids)
    add("attention B B=128 N=197 H=12 ATS mask + 177 gathered query rows", mk_attn_ats)

    def mk_attn_scores(b=128, h=12, n=197):
        qkv = rnd(b, n, 3 * h * 64, dtype=torch.bfloat16)
        mask = torch.ones(b, n, dtype=torch.bool, device=DEV)
        return lambda: T.attention(qkv, h, 0.125, None, mask, None, False, True, False)
    add("attention B B=128 N=197 H=12 scores only (CLS rows, v not read)", mk_attn_scores)
    # residual add + LayerNorm + bf16 cast
    for b, n, c in ((256, 197, 384), (256, 97, 384), (128, 197, 768), (1024, 197, 768)):
        def mk_ln(b=b, n=n, c=c, br=True):
            x, w, bb = rnd(b, n, c), rnd(c), rnd(c)
            branch = rnd(b, n, c, dtype=torch.bfloat16) if br else None
            return lambda: T.add_layernorm(x, branch, w, bb, 1e-6)
        add(f"add_layernorm B={b} N={n} C={c} (x + branch, LN, bf16)", mk_ln)
        if b == 256 and n == 197:
            add(f"add_layernorm B={b} N={n} C={c} (LN, bf16; no branch)", lambda f=mk_ln: f(br=False))
    # in front of block 0: image -> bf16 patch rows; cat(cls, patches) + pos_embed + the first norm1
    for b, c in ((256, 384), (128, 768), (1024, 768)):
        def mk_patch(b=b):
            img = rnd(b, 3, 224, 224)
            return lambda: T.patchify(img, 16, 16)

        def mk_embed(b=b, c=c):
            patches, tok, pos, w, bb = rnd(b, 196, c, dtype=torch.bfloat16), rnd(1, c), rnd(197, c), rnd(c), rnd(c)
            return lambda: T.embed_layernorm(patches, tok, pos, w, bb, 1e-6)
        if c != 384 or b == 256:
            add(f"patchify B={b} 3x224x224 -> [196, 768] bf16", mk_patch)
        add(f"embed_layernorm B={b} N=197 C={c} (cat cls + pos_embed + norm1, bf16)", mk_embed)
    # the residual add on the rows a select kernel reads (bf16-autocast Top-K / EViT blocks)
    for b, n, k in ((128, 197, 98), (1024, 197, 98)):
        def mk_topk_add(b=b, n=n, k=k):
            x, br, s = rnd(b, n, 768), rnd(b, n, 768, dtype=torch.bfloat16), spread_scores(b, n - 1, 1).to(DEV)
            return lambda: T.topk_gather_add(x, br, s, k)

        def mk_evit_add(b=b, n=n, k=k):
            x, br, s = rnd(b, n, 768), rnd(b, n, 768, dtype=torch.bfloat16), spread_scores(b, n - 1, 1).to(DEV)
            return lambda: T.evit_select_fuse_add(x, br, s, k)
        add(f"topk_gather_add B B={b} N={n} k={k} (x + attn branch on the kept rows)", mk_topk_add)
        add(f"evit_select_fuse_add B B={b} N={n} k={k} (x + attn branch on every row read)", mk_evit_add)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "ops_r02.json"))
    ap.add_argument("--launches", type=int, default=24, help="launches per event pair")
    ap.add_argument("--reps", type=int, default=5, help="event pairs per case (median)")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    peaks, kind = measured_peaks()
    l2_cycle = 320e6            # bytes that must pass between two uses of the same replica (> 2 x 126 MB L2)
    rows = []
    for label, make in cases():
        if a.only and a.only not in label:
            continue
        # first replica: learn the launch's name / algorithmic bytes
        thunks = [make()]
        _lib.TIMELINE = []
        keep = thunks[0]()
        tl, _lib.TIMELINE = _lib.TIMELINE, None
        assert len(tl) == 1, f"{label}: expected exactly one launch, got {len(tl)}"
        kname, kargs = tl[0][0], tl[0][1]
        nbytes = algorithmic_bytes(kname, kargs)
        free = torch.cuda.mem_get_info()[0]
        used_one = max(torch.cuda.memory_allocated(), 1)
        reps_needed = max(2, int(l2_cycle // max(nbytes, 1)) + 1)
        n_rep = int(min(reps_needed, 48, max(2, (free * 0.6) // max(nbytes * 3, 1))))
        while len(thunks) < n_rep:
            thunks.append(make())
        ring = [None] * n_rep           # keeps the last outputs alive: the allocator cannot hand back the same block
        for i in range(n_rep):
            ring[i] = thunks[i]()
        torch.cuda.synchronize()
        per_launch = []
        for _ in range(a.reps):
            torch.cuda._sleep(3_000_000)         # ~1.5 ms of device time: the host enqueues the whole batch behind it
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.launches):
                ring[i % n_rep] = thunks[i % n_rep]()
            e1.record()
            torch.cuda.synchronize()
            per_launch.append(e0.elapsed_time(e1) * 1e3 / a.launches)
        per_launch.sort()
        med = per_launch[len(per_launch) // 2]
        gbs = nbytes / (med * 1e-6) / 1e9
        resident = n_rep * nbytes < 2 * 126e6
        rows.append({"case": label, "kernel": kname.replace("tokred_", ""), "us": round(med, 2), "us_min": round(per_launch[0], 2),
                     "alg_mb": round(nbytes / 1e6, 3), "alg_gbs": round(gbs, 1), "frac_hbm": round(gbs / peaks["hbm_gbs"], 4),
                     "roofline_us": round(nbytes / (peaks["hbm_gbs"] * 1e9) * 1e6, 2), "replicas": n_rep,
                     "l2_resident": bool(resident)})
        print(f"{label:62s} {med:8.2f} us  {nbytes / 1e6:9.2f} MB  {gbs:8.1f} GB/s  {100 * gbs / peaks['hbm_gbs']:5.1f}% of {kind} HBM peak"
              f"  x{n_rep}{' (L2-resident)' if resident else ''}", flush=True)
        del thunks, ring, keep
        torch.cuda.empty_cache()
    with open(a.out, "w") as fh:
        json.dump({"peak_hbm_gbs": peaks["hbm_gbs"], "peak_source": kind, "gpu": torch.cuda.get_device_name(0),
                   "timing": f"{a.launches} back-to-back launches per CUDA-event pair (queued behind a device-side sleep so the "
                             "bracket holds no host latency), cycling input replicas totalling > 2x L2 with their outputs kept "
                             f"alive; median of {a.reps} pairs; per-launch = elapsed / {a.launches}", "rows": rows}, fh, indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
