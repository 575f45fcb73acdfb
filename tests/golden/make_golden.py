"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference, imported through
oracle/timm_shim.py) on seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

ops.pt     op-level vectors: inputs + outputs of every reference reduction function/module (fp32, CPU).
models.pt  model-level vectors: a micro DeiT (dim 32, 2 heads, depth 4, reduction at blocks 1/2/3, 196 patches)
           per method: state_dict, shared input images (fp16-stored), logits and per-stage decisions.
The reference ships no golden vectors or tests of its own, so these are the pin.
"""
import argparse
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import timm_shim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def g(seed):
    return torch.Generator().manual_seed(seed)


def rand_attn(b, h, n, seed, sharp=4.0):
    return torch.softmax(sharp * torch.randn(b, h, n, n, generator=g(seed)), dim=-1)


def make_ops(ref):
    out = {}
    b, n, c, h = 2, 65, 32, 2
    # --- Top-K / EViT (models/topk.py:55-65,89-93; models/evit.py:111-123)
    attn, x = rand_attn(b, h, n, 1), torch.randn(b, n, c, generator=g(2))
    k = 40
    cls_attn = attn[:, :, 0, 1:].mean(dim=1)
    _, idx = torch.topk(cls_attn, k, dim=1, largest=True, sorted=True)
    x_topk = torch.cat([x[:, :1], torch.gather(x[:, 1:], 1, idx.unsqueeze(-1).expand(-1, -1, c))], 1)
    compl = ref["evit"].complement_idx(idx, n - 1)
    extra = torch.sum(torch.gather(x[:, 1:], 1, compl.unsqueeze(-1).expand(-1, -1, c))
                      * torch.gather(cls_attn, 1, compl).unsqueeze(-1), dim=1, keepdim=True)
    out["select"] = dict(attn=attn, x=x, k=k, cls_attn=cls_attn, idx=idx, x_topk=x_topk, compl=compl,
                         x_evit=torch.cat([x_topk, extra], 1))
    # --- ToMe (models/tome.py:230-337)
    T = ref["tome"]
    nt, r = 197, 59
    metric, xt = torch.randn(b, nt, 16, generator=g(3)), torch.randn(b, nt, c, generator=g(4))
    size = torch.randint(1, 4, (b, nt, 1), generator=g(5)).float()
    merge, _ = T.bipartite_soft_matching(metric, r, True, False)
    xm, sm = T.merge_wavg(merge, xt, size)
    source = T.merge_source(merge, xt, None)
    rci = source * ((torch.ones(source.shape).permute(0, 2, 1)) * torch.arange(1, source.shape[1] + 1)).permute(0, 2, 1)
    rci = (torch.amax(rci, dim=-2) - 2)[:, 1:]
    out["tome"] = dict(metric=metric, x=xt, size=size, r=r, x_out=xm, size_out=sm, rci=rci)
    # --- DPC-KNN (models/dpcknn.py:44-140)
    D = ref["dpcknn"]
    p, kc = 64, 16
    xd = torch.randn(b, p, c, generator=g(6))
    torch.manual_seed(7)
    idx_cluster, index_down = D.cluster_dpc_knn(xd, kc, 5)
    torch.manual_seed(7)
    noise = torch.rand(b, p)
    tw = torch.randn(b, p, 1, generator=g(8)).exp()
    idx_token = torch.randint(0, p, (b, 196), generator=g(9))
    agg = torch.rand(b, 196, 1, generator=g(10))
    xmg, itn, awn = D.merge_tokens(xd, idx_token, agg, idx_cluster, kc, tw)
    out["dpcknn"] = dict(x=xd, K=kc, knn=5, noise=noise, dist=torch.cdist(xd, xd), idx_cluster=idx_cluster,
                         index_down=index_down, token_weight=tw, idx_token=idx_token, agg_weight=agg, x_merged=xmg,
                         idx_token_new=itn, agg_weight_new=awn)
    # --- K-Medoids (models/kmedoids.py:40-85,240)
    K = ref["kmedoids"]
    attn_k = rand_attn(b, h, p + 1, 11)
    twk = torch.sum(torch.sum(attn_k, dim=1), dim=1)[:, 1:].unsqueeze(2)
    cen, cidx, asg = K.k_medoids_fit(xd, kc, 3, twk)
    out["kmedoids"] = dict(x=xd, attn=attn_k, K=kc, iters=3, token_weight=twk, dist=torch.cdist(xd, xd), centres=cen,
                           cluster_idx=cidx, assignment=asg)
    # --- Sinkhorn / PatchMerger / SiT
    S = ref["sinkhorn"].Sinkhorn(c, 40, 1.0, 3)
    v0 = S.v.detach().clone()
    with torch.no_grad():
        xs, ws = S(xd)
    out["sinkhorn"] = dict(x=xd, v=v0, eps=1.0, iters=3, out=xs, weights=ws, v_after=S.v.detach().clone())
    PM = ref["patchmerger"].PatchMerger(c, 40)
    with torch.no_grad():
        PM.norm.weight.copy_(torch.rand(c, generator=g(12)) + 0.5)
        PM.norm.bias.copy_(torch.randn(c, generator=g(13)) * 0.1)
        xp, ap = PM(xd)
    out["patchmerger"] = dict(x=xd, ln_w=PM.norm.weight.detach().clone(), ln_b=PM.norm.bias.detach().clone(),
                              queries=PM.queries.detach().clone(), out=xp, attn=ap)
    ST = ref["sit"].TokenSlimmingModule(c, 40)
    with torch.no_grad():
        ST.scale.fill_(1.3)
        xs2, ws2 = ST(xd)
        logits = ST.weight(xd)
    out["sit"] = dict(x=xd, logits=logits, scale=ST.scale.detach().clone(), out=xs2, weights=ws2)
    # --- ATS (models/ats.py:44-89)
    A = ref["ats"].AdaptiveTokenSampling(41)
    attn_a = rand_attn(b, h, n, 14, sharp=6.0)
    va = torch.randn(b, h, n, 16, generator=g(15))
    mask = torch.ones(b, n, dtype=torch.bool)
    mask[1, n - 9:] = False
    na, nm, ids = A(va, attn_a, mask)
    out["ats"] = dict(v=va, attn=attn_a, mask=mask, sample_count=41, new_attn=na, new_mask=nm, ids=ids)
    # --- DynamicViT (models/dyvit.py:113-119,231-236,340-356)
    Dy = ref["dyvit"]
    P = Dy.PredictorLG(c).eval()
    policy = (torch.rand(b, p, 1, generator=g(16)) > 0.3).float()
    with torch.no_grad():
        hh = P.in_conv(xd)
        local_x = hh[:, :, : c // 2]
        global_x = (hh[:, :, c // 2:] * policy).sum(dim=1, keepdim=True) / torch.sum(policy, dim=1, keepdim=True) + P.eps
        feat = torch.cat([local_x, global_x.expand(b, p, c // 2)], dim=-1)
    score = torch.randn(b, p, generator=g(17))
    keep = torch.argsort(score, dim=1, descending=True)[:, :30]
    xx = torch.randn(b, p + 1, c, generator=g(18))
    now = torch.cat([torch.zeros(b, 1, dtype=keep.dtype), keep + 1], dim=1)
    out["dyvit"] = dict(h=hh, policy=policy, feat=feat, score=score, k=30, x=xx, keep=keep,
                        x_out=Dy.batch_index_select(xx, now))
    return out


METHODS = {"topk": ("TopKVisionTransformer", 0.7), "evit": ("EfficientVisionTransformer", 0.5),
           "tome": ("ToMeVisionTransformer", 0.7), "dyvit": ("DynamicVisionTransformer", 0.5),
           "dpcknn": ("DPCKNNVisionTransformer", 0.25), "kmedoids": ("KMedoidsVisionTransformer", 0.25),
           "sinkhorn": ("SinkhornVisionTransformer", 0.9), "patchmerger": ("PatchMergerVisionTransformer", 0.9),
           "ats": ("ATSVisionTransformer", 0.9), "sit": ("SelfSlimmedVisionTransformer", 0.9)}
MICRO = dict(embed_dim=32, depth=4, num_heads=2, num_classes=10)
MICRO_LOC = [1, 2, 3]


def micro_args(kr):
    return argparse.Namespace(keep_rate=[kr], reduction_loc=list(MICRO_LOC), distillation_type="none", k_neighbors=5,
                              cluster_iters=3, sinkhorn_eps=1.0, equal_weight=False, dyvit_distill=False, viz_mode=True)


def make_models(ref):
    images = torch.randn(2, 3, 224, 224, generator=g(100)).half()
    out = {"images_fp16": images, "micro": dict(MICRO), "reduction_loc": list(MICRO_LOC), "methods": {}}
    for name, (cls_name, kr) in METHODS.items():
        torch.manual_seed(200)
        with contextlib.redirect_stdout(io.StringIO()):
            model = getattr(ref[name], cls_name)(args=micro_args(kr), **MICRO).eval()
        with torch.no_grad():
            # spread the reduction-specific parameters so that decisions are not all-tied at random init (A.10)
            for n_, p_ in model.named_parameters():
                if n_.startswith("cluster_layers") and p_.dim() >= 2 and "queries" not in n_ and not n_.endswith(".v"):
                    p_.mul_(20.0)
                if n_.startswith("score_predictor") and p_.dim() >= 2:
                    p_.mul_(4.0)      # larger gains saturate the log-softmax into exact ties
        sd = {k_: v_.detach().clone() for k_, v_ in model.state_dict().items()}
        torch.manual_seed(300)
        with torch.no_grad():
            logits, viz = model(images.float())
        dec = {}
        for key in ("Kept_Tokens", "Assignment_Maps"):
            if key in viz:
                dec[key] = {int(i): torch.as_tensor(v) for i, v in viz[key].items()}
        out["methods"][name] = dict(keep_rate=kr, state_dict=sd, logits=logits, decisions=dec)
        print(name, logits.shape, float(logits.abs().mean()), {k_: list(v_.keys()) for k_, v_ in dec.items()})
    return out


def main():
    timm_shim.import_reference()
    import importlib
    ref = {n: importlib.import_module(f"models.{n}") for n in
           ["topk", "evit", "tome", "dpcknn", "kmedoids", "sinkhorn", "ats", "dyvit", "patchmerger", "sit"]}
    torch.save(make_ops(ref), os.path.join(HERE, "ops.pt"))
    torch.save(make_models(ref), os.path.join(HERE, "models.pt"))
    for f in ("ops.pt", "models.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
