"""TEST INFRASTRUCTURE ONLY — functional restatement of the reference's reduced-ViT forward passes.

``forward(method, state_dict, images, cfg)`` runs the per-method block loop of models/<method>.py (cited per
branch) over a plain state_dict with the reference's parameter names, calling the oracle operators in
oracle/ops.py at every reduction stage.  It is device-agnostic: on CPU it is the body of bench.py's cpu_baseline
and ``--impl reference`` legs; in the GPU tests it runs on the same device as the product model so that the
backbone (cuBLAS) numerics are identical and only the reduction operators differ.

Pinned against the unmodified reference models by tests/test_oracle_vs_reference.py::test_model_* (build
container) and by tests/golden/model_*.pt elsewhere.

``amp=True`` emulates ``torch.autocast(device, dtype=bfloat16)``: the backbone runs under autocast, the oracle
operators run with autocast disabled and ``lowp=torch.bfloat16`` (they round where CUDA autocast rounds).
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import ops as O

Tensor = torch.Tensor


@dataclass
class Cfg:
    embed_dim: int = 384
    num_heads: int = 6
    depth: int = 12
    patch_size: int = 16
    keep_rate: List[float] = field(default_factory=lambda: [0.7])
    reduction_loc: List[int] = field(default_factory=lambda: [3, 6, 9])
    k_neighbors: int = 5
    cluster_iters: int = 3
    sinkhorn_eps: float = 1.0
    num_patches: int = 196

    def rates(self):
        r = list(self.keep_rate)
        return [r[0] ** (i + 1) for i in range(len(self.reduction_loc))] if len(r) == 1 else r

    def counts(self):
        r = list(self.keep_rate)
        if len(r) == 1:
            return [int(self.num_patches * r[0] ** (i + 1)) for i in range(len(self.reduction_loc))]
        return [int(v) for v in r]


def _ln(x, sd, name, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _lin(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _embed(sd, images, cfg: Cfg):
    x = F.conv2d(images, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=cfg.patch_size)
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat((sd["cls_token"].expand(x.shape[0], -1, -1), x), dim=1)
    return x + sd["pos_embed"]


def _qkv(x, sd, i, heads):
    b, n, c = x.shape
    qkv = _lin(x, sd, f"blocks.{i}.attn.qkv").reshape(b, n, 3, heads, c // heads).permute(2, 0, 3, 1, 4)
    return qkv[0], qkv[1], qkv[2]


def _attend(x, sd, i, heads, size=None, mask=None):
    """norm1 -> qkv -> softmax probabilities.  Returns (attn, k, v)."""
    q, k, v = _qkv(_ln(x, sd, f"blocks.{i}.norm1"), sd, i, heads)
    dots = (q @ k.transpose(-2, -1)) * ((x.shape[-1] // heads) ** -0.5)
    if size is not None:
        dots = dots + size.log()[:, None, None, :, 0]
    if mask is not None:
        m2 = mask.unsqueeze(1).unsqueeze(3) * mask.unsqueeze(1).unsqueeze(2)
        dots = dots.masked_fill(~m2, -torch.finfo(dots.dtype).max)
    return dots.softmax(dim=-1), k, v


def _proj(attn, v, sd, i):
    b, _, n, _ = attn.shape
    return _lin((attn @ v).transpose(1, 2).reshape(b, n, -1), sd, f"blocks.{i}.attn.proj")


def _mlp(x, sd, i):
    h = _ln(x, sd, f"blocks.{i}.norm2")
    return x + _lin(F.gelu(_lin(h, sd, f"blocks.{i}.mlp.fc1")), sd, f"blocks.{i}.mlp.fc2")


def _plain_block(x, sd, i, heads):
    attn, _, v = _attend(x, sd, i, heads)
    return _mlp(x + _proj(attn, v, sd, i), sd, i), attn


def _head(x, sd):
    x = _ln(x, sd, "norm")
    return _lin(x[:, 0], sd, "head")


@torch.no_grad()
def forward(method: str, sd: Dict[str, Tensor], images: Tensor, cfg: Cfg, amp: bool = False,
            record: Optional[dict] = None) -> Tensor:
    dev = images.device.type
    lowp = torch.bfloat16 if amp else None
    ac = (lambda: torch.autocast(dev, dtype=torch.bfloat16)) if amp else contextlib.nullcontext
    noac = lambda: torch.autocast(dev, enabled=False)
    rec = record if record is not None else {}
    heads, loc = cfg.num_heads, list(cfg.reduction_loc)
    with ac():
        x = _embed(sd, images, cfg)
        b = x.shape[0]
        stage = 0

        if method in ("topk", "evit"):                          # models/topk.py:179-203, models/evit.py:111-123
            rates = cfg.rates()
            for i in range(cfg.depth):
                attn, _, v = _attend(x, sd, i, heads)
                x = x + _proj(attn, v, sd, i)
                if i in loc:
                    k = int(rates[loc.index(i)] * 196)
                    if k != x.shape[1] - 1:
                        with noac():
                            scores = O.cls_attention_scores(attn)
                            if method == "topk":
                                x, idx = O.topk_gather(x, scores, k)
                            else:
                                x, idx, compl = O.evit_select_fuse(x, scores, k)
                        rec[i] = idx
                        rec[("in", i)] = {"scores": scores, "k": k}
                x = _mlp(x, sd, i)

        elif method == "tome":                                  # models/tome.py:78-104, 202-203
            counts = cfg.counts()
            size = None
            prev = cfg.num_patches
            r_at = {}
            for j, l in enumerate(loc):
                r_at[l] = prev - counts[j]
                prev = counts[j]
            for i in range(cfg.depth):
                attn, k, v = _attend(x, sd, i, heads, size=size)
                x = x + _proj(attn, v, sd, i)
                if r_at.get(i, 0) > 0 and O.tome_effective_r(x.shape[1], r_at[i]) > 0:
                    with noac():
                        metric = k.mean(1)
                        unm, src, dst, _ = O.tome_match(metric, r_at[i], True, lowp=lowp)
                        x, size, rci = O.tome_merge(x, size, unm, src, dst)
                    rec[i] = rci
                    rec[("in", i)] = {"metric": metric, "r": r_at[i]}
                x = _mlp(x, sd, i)

        elif method == "dyvit":                                 # models/dyvit.py:205-238
            rates = cfg.rates()
            prev_decision = torch.ones(b, cfg.num_patches, 1, dtype=x.dtype, device=x.device)
            for i in range(cfg.depth):
                if i in loc:
                    j = loc.index(i)
                    pre = f"score_predictor.{j}"
                    h = F.gelu(_lin(_ln(x[:, 1:], sd, pre + ".in_conv.0", 1e-5), sd, pre + ".in_conv.1"))
                    with noac():
                        feat = O.dyvit_pool_concat(h, prev_decision)
                    h = F.gelu(_lin(feat, sd, pre + ".out_conv.0"))
                    h = F.gelu(_lin(h, sd, pre + ".out_conv.2"))
                    score = F.log_softmax(_lin(h, sd, pre + ".out_conv.4"), dim=-1)[:, :, 0]
                    k = int(cfg.num_patches * rates[j])
                    with noac():
                        x, keep = O.dyvit_keep(x, score, k)
                    prev_decision = torch.gather(prev_decision, 1, keep.unsqueeze(-1))
                    rec[i] = keep
                    rec[("in", i)] = {"scores": score.float(), "k": k}
                x, _ = _plain_block(x, sd, i, heads)

        elif method == "dpcknn":                                # models/dpcknn.py:231-268
            counts = cfg.counts()
            idx_token = torch.arange(cfg.num_patches, device=x.device)[None, :].repeat(b, 1)
            agg_weight = x.new_ones(b, cfg.num_patches, 1)
            for i in range(cfg.depth):
                if i in loc:
                    j = loc.index(i)
                    cls, xp = x[:, :1], x[:, 1:]
                    tw = _lin(xp, sd, f"cluster_layers.{j}.score").exp()
                    with noac():
                        noise = torch.rand((b, xp.shape[1]), device=x.device, dtype=torch.float32)
                        idx_cluster, index_down = O.dpcknn_cluster(xp.float(), counts[j], cfg.k_neighbors, noise)
                        xm, idx_token, agg_weight = O.dpcknn_merge(xp, idx_token, agg_weight, idx_cluster, counts[j], tw.float())
                    rec[i] = (index_down, idx_cluster)
                    rec[("in", i)] = {"x": xp.float(), "noise": noise, "K": counts[j], "knn": cfg.k_neighbors}
                    x = torch.cat((cls, xm), dim=1)
                x, _ = _plain_block(x, sd, i, heads)

        elif method == "kmedoids":                              # models/kmedoids.py:225-251
            counts = cfg.counts()
            attn = None
            for i in range(cfg.depth):
                if i in loc:
                    j = loc.index(i)
                    cls, xp = x[:, :1], x[:, 1:]
                    with noac():
                        tw = O.attn_colsum(attn.float())
                        centres, cidx, assign = O.kmedoids_fit(xp.float(), counts[j], cfg.cluster_iters, tw)
                    rec[i] = (cidx, assign)
                    rec[("in", i)] = {"x": xp.float(), "tw": tw, "K": counts[j], "iters": cfg.cluster_iters}
                    x = torch.cat((cls, centres.to(x.dtype)), dim=1)
                x, attn = _plain_block(x, sd, i, heads)

        elif method in ("sinkhorn", "patchmerger", "sit"):      # models/sinkhorn.py:164-182 and siblings
            counts = cfg.counts()
            for i in range(cfg.depth):
                if i in loc:
                    j = loc.index(i)
                    cls, xp = x[:, :1], x[:, 1:]
                    pre = f"cluster_layers.{j}"
                    if method == "sit":
                        hdn = F.gelu(_lin(_ln(xp, sd, pre + ".weight.0", 1e-5), sd, pre + ".weight.1"))
                        logits = _lin(hdn, sd, pre + ".weight.3")
                    with noac():
                        if method == "sinkhorn":
                            xm, w, vh = O.sinkhorn_merge(xp, sd[pre + ".v"], cfg.sinkhorn_eps, cfg.cluster_iters, lowp=lowp)
                            sd[pre + ".v"].copy_(vh)                      # the reference overwrites its parameter
                        elif method == "patchmerger":
                            xm, w = O.patchmerger(xp, sd[pre + ".norm.weight"], sd[pre + ".norm.bias"], sd[pre + ".queries"], lowp=lowp)
                        else:
                            xm, w = O.sit_merge(xp, logits, sd[pre + ".scale"], lowp=lowp)
                    rec[i] = w
                    x = torch.cat((cls, xm.to(cls.dtype)), dim=1)
                x, _ = _plain_block(x, sd, i, heads)

        elif method == "ats":                                   # models/ats.py:110-134,152-162,228-246
            counts = cfg.counts()
            sample_count = {l: counts[j] + 1 for j, l in enumerate(loc)} if len(cfg.keep_rate) == 1 else \
                {l: int(cfg.keep_rate[j]) for j, l in enumerate(loc)}
            mask = torch.ones((b, x.shape[1]), dtype=torch.bool, device=x.device)
            for i in range(cfg.depth):
                attn, _, v = _attend(x, sd, i, heads, mask=mask)
                if i in sample_count:
                    with noac():
                        rec[("in", i)] = {"v": v, "attn": attn.float(), "mask": mask, "count": sample_count[i]}
                        attn, mask, ids = O.ats_sample(v, attn.float(), mask, sample_count[i])
                    x = O.gather_rows(x, ids)
                    rec[i] = ids
                x = x + _proj(attn, v, sd, i)
                x = _mlp(x, sd, i)
        else:
            raise ValueError(method)
        return _head(x, sd)


def cfg_for(size: str, **kw) -> Cfg:
    dims = {"tiny": (192, 3), "small": (384, 6), "base": (768, 12)}[size]
    return Cfg(embed_dim=dims[0], num_heads=dims[1], **kw)


# ------------------------------------------------------------------------------------------------ f4: training paths
def softmax_with_policy(attn: Tensor, policy: Tensor, eps: float = 1e-6) -> Tensor:
    """models/dyvit.py:39-51."""
    b, n, _ = policy.shape
    pol = policy.reshape(b, 1, 1, n)
    eye = torch.eye(n, dtype=pol.dtype, device=pol.device).view(1, 1, n, n)
    pol = pol + (1.0 - pol) * eye
    mx = attn.max(dim=-1, keepdim=True)[0]
    a = (attn - mx).to(torch.float32).exp() * pol.to(torch.float32)
    a = (a + eps / n) / (a.sum(dim=-1, keepdim=True) + eps)
    return a.type_as(mx)


def dyvit_train_forward(sd: Dict[str, Tensor], images: Tensor, cfg: Cfg):
    """DynamicViT training forward, models/dyvit.py:205-229 + :251-261 (dyvit_distillation=False): differentiable with
    respect to every tensor in ``sd`` that requires grad.  Returns (logits, [hard keep decisions per stage])."""
    heads, loc = cfg.num_heads, list(cfg.reduction_loc)
    x = _embed(sd, images, cfg)
    b = x.shape[0]
    n0 = cfg.num_patches
    prev = torch.ones(b, n0, 1, dtype=x.dtype, device=x.device)
    policy = torch.ones(b, n0 + 1, 1, dtype=x.dtype, device=x.device)
    out_pred = []

    def block(x, i, policy):
        q, k, v = _qkv(_ln(x, sd, f"blocks.{i}.norm1"), sd, i, heads)
        dots = (q @ k.transpose(-2, -1)) * ((x.shape[-1] // heads) ** -0.5)
        attn = softmax_with_policy(dots, policy)
        return _mlp(x + _proj(attn, v, sd, i), sd, i)

    for i in range(cfg.depth):
        if i in loc:
            j = loc.index(i)
            pre = f"score_predictor.{j}"
            h = F.gelu(_lin(_ln(x[:, 1:], sd, pre + ".in_conv.0", 1e-5), sd, pre + ".in_conv.1"))
            feat = O.dyvit_pool_concat(h, prev)
            h = F.gelu(_lin(feat, sd, pre + ".out_conv.0"))
            h = F.gelu(_lin(h, sd, pre + ".out_conv.2"))
            score = F.log_softmax(_lin(h, sd, pre + ".out_conv.4"), dim=-1).reshape(b, -1, 2)
            hard = F.gumbel_softmax(score, hard=True)[:, :, 0:1] * prev                 # :216
            out_pred.append(hard.reshape(b, n0))
            policy = torch.cat([torch.ones(b, 1, 1, dtype=hard.dtype, device=hard.device), hard], dim=1)
            x = block(x, i, policy)
            prev = hard
        else:
            x = block(x, i, policy)
    return _head(x, sd), out_pred
