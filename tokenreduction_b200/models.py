"""Reduced-ViT model classes: the per-method ``*VisionTransformer`` forward loops of the reference
(models/<method>.py) over the tokred modules.  Constructor signatures, ``keep_rate`` / ``reduction_loc``
semantics, parameter names, ``get_new_module_names()`` / ``get_reduction_count()`` and the eval return values
(logits, or ``(logits, viz_dict)`` under ``viz_mode``) are the reference's.  The backbone (patch embedding,
attention, MLP, head) stays on PyTorch/cuBLAS like the reference's.
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn

from . import modules as M
from . import ops
from .vit import PatchEmbed, VisionTransformer


class _Pending:
    """a viz tensor that stays on the device until the forward is over (the reference does one
    blocking ``.cpu().numpy()`` per stage and dict entry, e.g. models/topk.py:195-197)."""
    __slots__ = ("t",)

    def __init__(self, t):
        self.t = t


def _np(t):
    return _Pending(t.detach())


def _record_features(model, features, i, x):
    """``features[i] = x`` of the reference's viz dicts (e.g. models/topk.py:197).  ``model.viz_features = False`` (a
    multi-GPU caller that gathers only the kept / assignment indices, bench.py) skips the entry -- and with it the
    materialisation of a deferred residual sum that nothing else needs."""
    if getattr(model, "viz_features", True):
        features[i] = _np(M.value(x))


def _finalize_viz(viz, keep_on_device=False):
    """every pending tensor -> pinned host memory with async copies, ONE stream synchronisation, then numpy arrays
    (the types validate.py:199-229 consumes).  keep_on_device (args.tokred_device_viz): hand the device tensors out
    instead -- no copy, no synchronisation; a multi-GPU caller all_gathers them (bench.py)."""
    if keep_on_device:
        def unwrap(d):
            for k, v in d.items():
                if isinstance(v, dict):
                    unwrap(v)
                elif isinstance(v, _Pending):
                    d[k] = v.t
        unwrap(viz)
        return viz
    pend = []

    def walk(d):
        for v in d.values():
            if isinstance(v, dict):
                walk(v)
            elif isinstance(v, _Pending):
                pend.append(v)
    walk(viz)
    host = []
    for p_ in pend:
        if p_.t.is_cuda:
            h = torch.empty(p_.t.shape, dtype=p_.t.dtype, pin_memory=True)
            h.copy_(p_.t, non_blocking=True)
        else:
            h = p_.t.clone()
        host.append(h)
    if any(p_.t.is_cuda for p_ in pend):
        torch.cuda.current_stream().synchronize()
    done = {id(p_): h.numpy() for p_, h in zip(pend, host)}

    def swap(d):
        for k, v in d.items():
            if isinstance(v, dict):
                swap(v)
            elif isinstance(v, _Pending):
                d[k] = done[id(v)]
    swap(viz)
    return viz


def _geometric(keep_rate, n_stages, what):
    """a single keep rate r expands to r, r^2, r^3 ... (e.g. models/topk.py:141-142)."""
    rates = list(keep_rate)
    if len(rates) == 1:
        rates = [rates[0] ** (i + 1) for i in range(n_stages)]
    assert len(rates) == n_stages, f"Mismatch between the {what} location and token ratios ({rates})"
    return rates


def _counts(keep_rate, n_stages, num_patches, what):
    """merging / clustering methods use absolute counts int(196 * r^(i+1)) (e.g. models/tome.py:145-146)."""
    counts = list(keep_rate)
    if len(counts) == 1:
        counts = [int(num_patches * counts[0] ** (i + 1)) for i in range(n_stages)]
    assert len(counts) == n_stages, f"Mismatch between the {what} location and cluster counts ({counts})"
    return counts


class _ReducedViT(VisionTransformer):
    """shared head/tail of every reduced model's forward."""

    def _logits(self, x):
        # LayerNorm is per token: norm(x)[:, 0] == norm(x[:, 0]) bit for bit, and only the class token is read
        # (e.g. models/topk.py:205-207)
        return self.head(self.pre_logits(self.norm(M.cls_value(x))))

    def _ret(self, x, viz_data=None):
        if self.training or not self.viz_mode:
            return x
        return x, _finalize_viz(viz_data, getattr(self, "device_viz", False))


# =============================================================================================== Top-K / EViT
class TopKVisionTransformer(_ReducedViT):
    """models/topk.py:102-212."""
    block_cls = M.Block_TopK

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, representation_size=None, distilled=False, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.0, embed_layer=PatchEmbed, norm_layer=None, act_layer=None,
                 weight_init="", args=None, dyvit_distillation=False):
        super().__init__(img_size, patch_size, in_chans, num_classes, embed_dim, depth, num_heads, mlp_ratio, qkv_bias,
                         representation_size, distilled, drop_rate, attn_drop_rate, drop_path_rate, embed_layer,
                         norm_layer, act_layer, weight_init)
        pruning_loc = args.reduction_loc
        token_ratio = _geometric(args.keep_rate, len(pruning_loc), "pruning")
        print(token_ratio, pruning_loc)
        ratio_full = [1 for _ in range(depth)]
        for i, loc in enumerate(pruning_loc):
            ratio_full[loc] = token_ratio[i]
        del self.blocks
        self.num_patches = self.patch_embed.num_patches
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        act_layer = act_layer or nn.GELU
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            self.block_cls(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop_rate,
                           attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer, keep_rate=ratio_full[i])
            for i in range(depth)])
        self.deit_distillation = distilled
        self.pruning_loc = pruning_loc
        self.token_ratio = token_ratio
        self.viz_mode = getattr(args, "viz_mode", False)
        self.device_viz = bool(getattr(args, "tokred_device_viz", False))
        self.apply(self._init_weights)

    def get_new_module_names(self):
        return []

    def get_reduction_count(self):
        return self.pruning_loc

    def forward(self, x):
        x = self.embed(x)
        decisions, features = {}, {}
        i = -1
        for i, blk in enumerate(self.blocks):
            out = blk(x)
            x, sample_idx = out[0], out[2]
            if self.viz_mode and sample_idx is not None:
                decisions[i] = _np(sample_idx)
                _record_features(self, features, i, x)
        if self.viz_mode and 11 not in features:
            _record_features(self, features, i, x)
        return self._ret(self._logits(x), {"Kept_Tokens": decisions, "Features": features})


class EfficientVisionTransformer(TopKVisionTransformer):
    """models/evit.py:132-244 — same loop over Block_EVIT (4-tuple per block)."""
    block_cls = M.Block_EVIT


# =============================================================================================== ToMe
class ToMeVisionTransformer(_ReducedViT):
    """models/tome.py:107-222."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, representation_size=None, distilled=False, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.0, embed_layer=PatchEmbed, norm_layer=None, act_layer=None,
                 weight_init="", args=None, dyvit_distillation=False):
        super().__init__(img_size, patch_size, in_chans, num_classes, embed_dim, depth, num_heads, mlp_ratio, qkv_bias,
                         representation_size, distilled, drop_rate, attn_drop_rate, drop_path_rate, embed_layer,
                         norm_layer, act_layer, weight_init)
        pruning_loc = args.reduction_loc
        token_ratio = _counts(args.keep_rate, len(pruning_loc), self.patch_embed.num_patches, "pruning")
        print(token_ratio, pruning_loc)
        r_full = [0 for _ in range(depth)]
        prev = self.patch_embed.num_patches
        for i, loc in enumerate(pruning_loc):        # counts -> per-stage r (models/tome.py:152-156)
            r_full[loc] = prev - token_ratio[i]
            prev = token_ratio[i]
        del self.blocks
        self.num_patches = self.patch_embed.num_patches
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        act_layer = act_layer or nn.GELU
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            M.Block_ToMe(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop_rate,
                         attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer, r=r_full[i])
            for i in range(depth)])
        self.deit_distillation = distilled
        self.pruning_loc = pruning_loc
        self.token_ratio = token_ratio
        self.prop_attn = True
        self.viz_mode = getattr(args, "viz_mode", False)
        self.device_viz = bool(getattr(args, "tokred_device_viz", False))
        self.apply(self._init_weights)

    def get_new_module_names(self):
        return []

    def get_reduction_count(self):
        return self.pruning_loc

    def forward(self, x):
        attn_size = None
        x = self.embed(x)
        assignments, features = {}, {}
        i = -1
        for i, blk in enumerate(self.blocks):
            x, attn_size, cluster_assign = blk(x, attn_size)
            if self.viz_mode and i in self.pruning_loc and cluster_assign is not None:
                assignments[i] = _np(cluster_assign)
                _record_features(self, features, i, x)
        if self.viz_mode and 11 not in features:
            _record_features(self, features, i, x)
        return self._ret(self._logits(x), {"Assignment_Maps": assignments, "Features": features})


# =============================================================================================== cluster-layer models
class _ClusterLayerViT(_ReducedViT):
    """DPC-KNN / Sinkhorn / PatchMerger / SiT share this shape: stock blocks + ``cluster_layers`` applied to the
    patch tokens before block i for i in cluster_loc (models/dpcknn.py:256-268 etc.)."""

    def _setup(self, args, make_layer):
        self.cluster_loc = args.reduction_loc
        self.cluster_count = _counts(args.keep_rate, len(self.cluster_loc), self.patch_embed.num_patches, "cluster")
        print(self.cluster_count, self.cluster_loc)
        self.cluster_layers = nn.ModuleList([make_layer(c) for c in self.cluster_count])
        self.blocks = nn.ModuleList([*self.blocks])
        self.viz_mode = getattr(args, "viz_mode", False)
        self.device_viz = bool(getattr(args, "tokred_device_viz", False))
        self.apply(self._init_weights)

    def get_new_module_names(self):
        return ["cluster_layers"]

    def get_reduction_count(self):
        return self.cluster_loc


class DPCKNNVisionTransformer(_ClusterLayerViT):
    """models/dpcknn.py:175-285."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, representation_size=None, distilled=False, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.0, embed_layer=PatchEmbed, norm_layer=None, act_layer=None,
                 weight_init="", args=None):
        super().__init__(img_size, patch_size, in_chans, num_classes, embed_dim, depth, num_heads, mlp_ratio, qkv_bias,
                         representation_size, distilled, drop_rate, attn_drop_rate, drop_path_rate, embed_layer,
                         norm_layer, act_layer, weight_init)
        self.k_neighbors = args.k_neighbors
        self.equal_weight = args.equal_weight
        self._setup(args, lambda c: M.CTM(embed_dim, c, self.k_neighbors, self.equal_weight))

    def forward(self, x):
        x = self.patch_embed(x)
        b, n, _ = x.shape
        idx_token = torch.arange(n, device=x.device)[None, :].repeat(b, 1)
        agg_weight = x.new_ones(b, n, 1)
        x = self.embed_tokens(x)
        cnt = 0
        decisions, assignments, centers_feats, features = {}, {}, {}, {}
        i = -1
        for i, blk in enumerate(self.blocks):
            if i in self.cluster_loc:
                x = M.value(x)
                global_tokens = x[:, :self.num_tokens]
                x, idx_token, agg_weight, idx_centers, idx_cluster, cluster_centers = self.cluster_layers[cnt](
                    x[:, self.num_tokens:], idx_token, agg_weight, self.viz_mode)
                x = M.concat_tokens(global_tokens, x, blk.norm1)
                cnt += 1
                if self.viz_mode:
                    decisions[i], assignments[i], centers_feats[i] = _np(idx_centers), _np(idx_cluster), _np(cluster_centers)
            x = blk(x)
            if self.viz_mode:
                _record_features(self, features, i, x)
        if self.viz_mode and 11 not in features:
            _record_features(self, features, i, x)
        return self._ret(self._logits(x), {"Kept_Tokens": decisions, "Assignment_Maps": assignments,
                                           "Center_Feats": centers_feats, "Features": features})


class KMedoidsVisionTransformer(_ClusterLayerViT):
    """models/kmedoids.py:151-268.  Blocks return (x, attn); the previous block's attention gives token weights."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, representation_size=None, distilled=False, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.0, embed_layer=PatchEmbed, norm_layer=None, act_layer=None,
                 weight_init="", args=None):
        super().__init__(img_size, patch_size, in_chans, num_classes, embed_dim, depth, num_heads, mlp_ratio, qkv_bias,
                         representation_size, distilled, drop_rate, attn_drop_rate, drop_path_rate, embed_layer,
                         norm_layer, act_layer, weight_init)
        self.cluster_iters = args.cluster_iters
        self.equal_weight = args.equal_weight
        del self.blocks
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        act_layer = act_layer or nn.GELU
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            M.BlockWithProbs(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop_rate,
                             attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer, act_layer=act_layer)
            for i in range(depth)])
        self.fused_token_weights = bool(getattr(args, "tokred_fused_scores", False))
        self._setup(args, lambda c: M.KMedoids(c, self.cluster_iters, self.equal_weight))
        # only the block in front of a cluster layer has its probabilities read, and only their column sums
        # (models/kmedoids.py:240): under bf16 autocast no block materialises [B,H,N,N] (modules.AttentionWithProbs)
        for i, blk in enumerate(self.blocks):
            blk.attn.probs_mode = "colsum" if (i + 1) in self.cluster_loc else "none"

    def forward(self, x):
        x = self.patch_embed(x)
        b = x.shape[0]
        x = self.embed_tokens(x)
        cnt = 0
        attn = None
        decisions, assignments, centers_feats, features = {}, {}, {}, {}
        i = -1
        for i, blk in enumerate(self.blocks):
            if i in self.cluster_loc:
                x = M.value(x)
                global_tokens = x[:, :self.num_tokens]
                if attn.dim() == 3:      # per-head column sums [B,H,N] from the fused attention (deterministic order)
                    token_weights = attn.sum(1)[:, self.num_tokens:].unsqueeze(2)
                elif self.fused_token_weights:
                    token_weights = ops.attn_colsum(attn, self.num_tokens)          # one pass over [B,H,N,N]
                else:   # the reference's two torch.sum launches: bit-identical decision input
                    token_weights = torch.sum(torch.sum(attn, dim=1), dim=1)[:, self.num_tokens:].unsqueeze(2)
                x, idx_centers, idx_cluster = self.cluster_layers[cnt](x[:, self.num_tokens:], token_weights)
                if self.viz_mode:
                    decisions[i], assignments[i], centers_feats[i] = _np(idx_centers), _np(idx_cluster), _np(x)
                x = M.concat_tokens(global_tokens, x, blk.norm1)
                cnt += 1
            x, attn = blk(x)
            if self.viz_mode:
                _record_features(self, features, i, x)
        if self.viz_mode and 11 not in features:
            _record_features(self, features, i, x)
        return self._ret(self._logits(x), {"Kept_Tokens": decisions, "Assignment_Maps": assignments,
                                           "Center_Feats": centers_feats, "Features": features})


class _SoftClusterViT(_ClusterLayerViT):
    """forward loop shared by Sinkhorn / PatchMerger / SiT (models/sinkhorn.py:164-182, patchmerger.py:115-133,
    sit.py:115-128): cluster layer returns (x, soft assignment [B,K,P])."""

    _has_center_feats = True

    def _layer_viz(self, cnt, b):
        return None

    def forward(self, x):
        x = self.patch_embed(x)
        b = x.shape[0]
        x = self.embed_tokens(x)
        cnt = 0
        assignments, hard_assignment, centers_feats, features = {}, {}, {}, {}
        i = -1
        for i, blk in enumerate(self.blocks):
            if i in self.cluster_loc:
                x = M.value(x)
                global_tokens = x[:, :self.num_tokens]
                x, soft_assign = self.cluster_layers[cnt](x[:, self.num_tokens:])
                x = M.concat_tokens(global_tokens, x, blk.norm1)
                if self.viz_mode:
                    assignments[i] = _np(soft_assign)
                    hard_assignment[i] = _np(torch.argmax(soft_assign, dim=-2))
                    cf = self._layer_viz(cnt, b)
                    if cf is not None:
                        centers_feats[i] = _np(cf)
                cnt += 1
            x = blk(x)
            if self.viz_mode:
                _record_features(self, features, i, x)
        if self.viz_mode and 11 not in features:
            _record_features(self, features, i, x)
        # key names of models/sinkhorn.py:197, models/patchmerger.py:148, models/sit.py:143
        viz = {"Assignment_Maps": hard_assignment, "Soft_Assignment_Maps": assignments, "Features": features}
        if self._has_center_feats:
            viz["Center_Feats"] = centers_feats
        return self._ret(self._logits(x), viz)


class SinkhornVisionTransformer(_SoftClusterViT):
    """models/sinkhorn.py:89-199."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, representation_size=None, distilled=False, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.0, embed_layer=PatchEmbed, norm_layer=None, act_layer=None,
                 weight_init="", args=None):
        super().__init__(img_size, patch_size, in_chans, num_classes, embed_dim, depth, num_heads, mlp_ratio, qkv_bias,
                         representation_size, distilled, drop_rate, attn_drop_rate, drop_path_rate, embed_layer,
                         norm_layer, act_layer, weight_init)
        self.sinkhorn_eps = args.sinkhorn_eps
        self.cluster_iters = args.cluster_iters
        self._setup(args, lambda c: M.Sinkhorn(embed_dim, c, self.sinkhorn_eps, self.cluster_iters))

    def _layer_viz(self, cnt, b):
        return self.cluster_layers[cnt].v.unsqueeze(0).expand(b, -1, -1)


class PatchMergerVisionTransformer(_SoftClusterViT):
    """models/patchmerger.py:42-150."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, representation_size=None, distilled=False, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.0, embed_layer=PatchEmbed, norm_layer=None, act_layer=None,
                 weight_init="", args=None):
        super().__init__(img_size, patch_size, in_chans, num_classes, embed_dim, depth, num_heads, mlp_ratio, qkv_bias,
                         representation_size, distilled, drop_rate, attn_drop_rate, drop_path_rate, embed_layer,
                         norm_layer, act_layer, weight_init)
        self._setup(args, lambda c: M.PatchMerger(embed_dim, c))

    def _layer_viz(self, cnt, b):
        return self.cluster_layers[cnt].queries.unsqueeze(0).expand(b, -1, -1)


class SelfSlimmedVisionTransformer(_SoftClusterViT):
    """models/sit.py:43-145."""
    _has_center_feats = False

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, representation_size=None, distilled=False, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.0, embed_layer=PatchEmbed, norm_layer=None, act_layer=None,
                 weight_init="", args=None):
        super().__init__(img_size, patch_size, in_chans, num_classes, embed_dim, depth, num_heads, mlp_ratio, qkv_bias,
                         representation_size, distilled, drop_rate, attn_drop_rate, drop_path_rate, embed_layer,
                         norm_layer, act_layer, weight_init)
        self._setup(args, lambda c: M.TokenSlimmingModule(embed_dim, c))


# =============================================================================================== ATS
class ATSVisionTransformer(_ReducedViT):
    """models/ats.py:165-271."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, representation_size=None, distilled=False, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.0, embed_layer=PatchEmbed, norm_layer=None, act_layer=None,
                 weight_init="", args=None):
        super().__init__(img_size, patch_size, in_chans, num_classes, embed_dim, depth, num_heads, mlp_ratio, qkv_bias,
                         representation_size, distilled, drop_rate, attn_drop_rate, drop_path_rate, embed_layer,
                         norm_layer, act_layer, weight_init)
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        act_layer = act_layer or nn.GELU
        self.sample_loc = args.reduction_loc
        sample_count = list(args.keep_rate)
        if len(sample_count) == 1:     # max sample count int(r^(i+1) * 196) + 1 (models/ats.py:204-205)
            sample_count = [int(args.keep_rate[0] ** (i + 1) * self.patch_embed.num_patches) + 1
                            for i in range(len(self.sample_loc))]
        assert len(sample_count) == len(self.sample_loc)
        cnt = 0
        self.sample_count = [0] * depth
        for i in range(depth):
            if i in self.sample_loc:
                self.sample_count[i] = sample_count[cnt]
                cnt += 1
        print(self.sample_count, self.sample_loc)
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            M.ATSBlock(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop_rate,
                       attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer, act_layer=act_layer,
                       ats_sample_count=self.sample_count[i])
            for i in range(depth)])
        # default: pad to sample_count -> no host read anywhere in the forward (north_star: "variable post-reduction token
        # counts are handled without host syncs"); args.tokred_ats_exact_width=True restores the reference's data-dependent
        # width max_b #unique (models/ats.py:77-83) at the cost of ONE scalar read per stage (the reference syncs per image)
        static = not bool(getattr(args, "tokred_ats_exact_width", False))
        for blk in self.blocks:
            if blk.attn.ats_sample_count:
                blk.attn.ats.static_width = static
        first = min([i for i, blk in enumerate(self.blocks) if blk.attn.ats_sample_count], default=len(self.blocks))
        for i, blk in enumerate(self.blocks):       # forward() hands these blocks the all-true mask it creates itself
            blk.attn.mask_all_true = i <= first
        self.viz_mode = getattr(args, "viz_mode", False)
        self.device_viz = bool(getattr(args, "tokred_device_viz", False))
        self.apply(self._init_weights)

    def get_new_module_names(self):
        return []

    def get_reduction_count(self):
        return self.sample_loc

    def forward(self, x):
        x = self.patch_embed(x)
        b, n = x.shape[:2]
        x = self.embed_tokens(x)
        mask = torch.ones((b, n + self.num_tokens), dtype=torch.bool, device=x.device)
        decisions, features = {}, {}
        i = -1
        for i, blk in enumerate(self.blocks):
            x, mask, sample_ids = blk(x, mask)
            if self.viz_mode and sample_ids is not None:
                decisions[i] = _np(sample_ids[:, 1:] - 1)
                _record_features(self, features, i, x)
        if self.viz_mode and 11 not in features:
            _record_features(self, features, i, x)
        return self._ret(self._logits(x), {"Kept_Tokens": decisions, "Features": features})


# =============================================================================================== DynamicViT
class DynamicVisionTransformer(_ReducedViT):
    """models/dyvit.py:122-268.  eval: keep top-k tokens (one select+gather launch per stage); training: gumbel-softmax
    keep decisions as a policy mask through softmax_with_policy (:205-229), differentiable end to end."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, representation_size=None, distilled=False,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, embed_layer=PatchEmbed, norm_layer=None,
                 act_layer=None, weight_init="", args=None, dyvit_distillation=False):
        super().__init__(img_size, patch_size, in_chans, num_classes, embed_dim, depth, num_heads, mlp_ratio, qkv_bias,
                         representation_size, distilled, drop_rate, attn_drop_rate, drop_path_rate, embed_layer,
                         norm_layer, act_layer, weight_init)
        pruning_loc = args.reduction_loc
        token_ratio = _geometric(args.keep_rate, len(pruning_loc), "pruning")
        print(token_ratio, pruning_loc)
        del self.blocks
        self.num_patches = self.patch_embed.num_patches
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        act_layer = act_layer or nn.GELU
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            M.Block_DyVIT(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                          drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer)
            for i in range(depth)])
        self.score_predictor = nn.ModuleList([M.PredictorLG(embed_dim) for _ in range(len(pruning_loc))])
        self.deit_distillation = distilled
        self.dyvit_distillation = dyvit_distillation
        self.pruning_loc = pruning_loc
        self.token_ratio = token_ratio
        self.viz_mode = getattr(args, "viz_mode", False)
        self.device_viz = bool(getattr(args, "tokred_device_viz", False))
        self.apply(self._init_weights)

    def get_new_module_names(self):
        return ["score_predictor"]

    def get_reduction_count(self):
        return self.pruning_loc

    def _forward_train(self, x):
        """models/dyvit.py:205-229,251-261: nothing is removed; the hard gumbel decisions mask the attention."""
        import torch.nn.functional as F
        b = x.shape[0]
        x = self.embed(x)
        p_count = 0
        out_pred_prob = []
        init_n = self.num_patches
        prev_decision = torch.ones(b, init_n, 1, dtype=x.dtype, device=x.device)
        policy = torch.ones(b, init_n + 1, 1, dtype=x.dtype, device=x.device)
        for i, blk in enumerate(self.blocks):
            if i in self.pruning_loc:
                x = M.value(x)
                pred_score = self.score_predictor[p_count](x[:, 1:], prev_decision).reshape(b, -1, 2)
                hard_keep_decision = F.gumbel_softmax(pred_score, hard=True)[:, :, 0:1] * prev_decision
                out_pred_prob.append(hard_keep_decision.reshape(b, init_n))
                cls_policy = torch.ones(b, 1, 1, dtype=hard_keep_decision.dtype, device=hard_keep_decision.device)
                policy = torch.cat([cls_policy, hard_keep_decision], dim=1)
                x = blk(x, policy=policy)
                prev_decision = hard_keep_decision
                p_count += 1
            else:
                x = blk(x, policy)
        x = self.norm(M.value(x))
        features = x[:, 1:]
        x = self.head(self.pre_logits(x[:, 0]))
        if self.dyvit_distillation:
            return x, features, prev_decision.detach(), out_pred_prob
        return x, out_pred_prob

    def forward(self, x):
        if self.training:
            return self._forward_train(x)
        b = x.shape[0]
        x = self.embed(x)
        p_count = 0
        init_n = self.num_patches
        prev_decision = torch.ones(b, init_n, 1, dtype=x.dtype, device=x.device)
        decisions, features_viz = {}, {}
        i = -1
        for i, blk in enumerate(self.blocks):
            if i in self.pruning_loc:
                x = M.value(x)
                pred_score = self.score_predictor[p_count](x[:, 1:], prev_decision).reshape(b, -1, 2)
                num_keep = int(init_n * self.token_ratio[p_count])
                # argsort(desc)[:k] + [CLS, keep+1] gather in one launch; scores read in place (stride 2)
                x, keep_policy = ops.topk_gather(x, pred_score[:, :, 0], num_keep)
                prev_decision = torch.gather(prev_decision, 1, keep_policy.unsqueeze(-1))
                x = blk(x)
                if self.viz_mode:
                    decisions[i] = _np(keep_policy)
                    _record_features(self, features_viz, i, x)
                p_count += 1
            else:
                x = blk(x)
        if self.viz_mode and 11 not in features_viz:
            _record_features(self, features_viz, i, x)
        return self._ret(self._logits(x), {"Kept_Tokens": decisions, "Features": features_viz})
