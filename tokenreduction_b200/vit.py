"""DeiT/ViT backbone with the timm-0.4.12 surface the reduced models subclass.

The reference subclasses ``timm.models.vision_transformer.VisionTransformer``
(timm==0.4.12, /root/reference/requirements.txt:8) and only touches a small
part of it: the constructor's positional order (models/topk.py:131-134), the
attributes ``patch_embed.num_patches / cls_token / dist_token / pos_embed /
pos_drop / blocks / norm / pre_logits / head / num_tokens`` and
``_init_weights``.  timm is not installed in this image, so the backbone is
written here from that description (models/deit_viz.py:75-212 documents the
base class).  Parameter names are identical to timm's so reference
checkpoints / state_dicts load unchanged.

The backbone is NOT the product: it stays on PyTorch/cuBLAS exactly as in the
reference.  The product is the reduction operators in ``csrc/``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from functools import partial
from typing import Callable, Dict

import torch
import torch.nn as nn

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


# --------------------------------------------------------------------------- init helpers
def trunc_normal_(tensor: torch.Tensor, mean: float = 0.0, std: float = 1.0, a: float = -2.0, b: float = 2.0):
    """Truncated normal in the absolute interval [a, b] (timm semantics)."""
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def lecun_normal_(tensor: torch.Tensor):
    fan_in = nn.init._calculate_fan_in_and_fan_out(tensor)[0]
    # variance-scaling, fan_in, truncated normal; .8796.. renormalises the truncation
    return trunc_normal_(tensor, std=math.sqrt(1.0 / fan_in) / 0.87962566103423978)


def _to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


# --------------------------------------------------------------------------- layers
class DropPath(nn.Module):
    """Per-sample stochastic depth. Identity in eval (the only mode on our path)."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob or 0.0)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x.div(keep) * mask


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None, flatten=True):
        super().__init__()
        self.img_size = _to_2tuple(img_size)
        self.patch_size = _to_2tuple(patch_size)
        self.grid_size = (self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        _, _, h, w = x.shape
        if (h, w) != self.img_size:
            raise AssertionError(f"Input image size ({h}*{w}) doesn't match model ({self.img_size[0]}*{self.img_size[1]}).")
        if self.flatten and self._as_gemm(x):
            # Non-overlapping patches: the stride-16 convolution IS a GEMM [B*196, 768] x [768, C].  Under bf16 autocast
            # cuDNN runs it as an implicit-GEMM convolution at 1.1 ms for B=256 (15 % of a ToMe DeiT-S step, whatever
            # cudnn.benchmark picks); as a cuBLAS GEMM on the unfolded patches it is ~0.1 ms with the same bf16 operands
            # and fp32 accumulation (accumulation order differs: a last-bit bf16 difference in a few elements).
            b, c, _, _ = x.shape
            gh, gw = self.grid_size
            ph, pw = self.patch_size
            if x.dtype == torch.float32 and pw % 4 == 0 and (c * ph * pw) % 8 == 0:
                from . import ops
                xb = ops.patchify(x, ph, pw)          # cast + patch-major permutation in one pass at HBM speed
            else:
                xb = x.to(torch.bfloat16).view(b, c, gh, ph, gw, pw).permute(0, 2, 4, 1, 3, 5).reshape(b, gh * gw, c * ph * pw)
            w = self.proj.weight.view(self.proj.weight.shape[0], -1)
            return self.norm(torch.nn.functional.linear(xb, w, self.proj.bias))
        x = self.proj(x)
        if self.flatten:
            x = x.flatten(2).transpose(1, 2)
        return self.norm(x)

    def _as_gemm(self, x) -> bool:
        from . import modules
        return (modules.FUSED_ATTENTION and x.is_cuda and not (self.training and torch.is_grad_enabled())
                and torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16
                and self.proj.stride == self.proj.kernel_size and self.proj.padding == (0, 0))


class Attention(nn.Module):
    """Plain multi-head self-attention with a materialised softmax, as in the reference
    (every reduced model needs the probabilities, e.g. models/topk.py:47-52)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def qkv_heads(self, x):
        b, n, c = x.shape
        qkv = self.qkv(x).reshape(b, n, 3, self.num_heads, c // self.num_heads).permute(2, 0, 3, 1, 4)
        return qkv[0], qkv[1], qkv[2]

    def attend(self, q, k):
        attn = (q @ k.transpose(-2, -1)) * self.scale
        return attn

    def project(self, attn, v):
        b, _, n, _ = attn.shape
        x = (attn @ v).transpose(1, 2).reshape(b, n, -1)
        return self.proj_drop(self.proj(x))

    def forward(self, x):
        from . import modules, ops
        if modules.fused_attention_ok(self, x, self.num_heads):        # no [B,H,N,N] tensor
            qkv = self.qkv(x)
            qkv = qkv if qkv.dtype == torch.bfloat16 else qkv.to(torch.bfloat16)
            return self.proj_drop(self.proj(ops.attention(qkv, self.num_heads, self.scale)[0]))
        q, k, v = self.qkv_heads(x)
        attn = self.attn_drop(self.attend(q, k).softmax(dim=-1))
        return self.project(attn, v)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x):
        from . import modules
        x, y = modules.enter_norm(self.norm1, x)       # x may be the previous block's deferred residual sum
        x, y = modules.add_norm(x, self.drop_path(self.attn(y)), self.norm2)
        return modules.defer_add(x, self.drop_path(self.mlp(y)), self.norm1, self)


def _init_vit_weights(module: nn.Module, name: str = "", head_bias: float = 0.0, jax_impl: bool = False):
    """timm-0.4.12 initialisation rule (described at models/deit_viz.py:215-246)."""
    if isinstance(module, nn.Linear):
        if name.startswith("head"):
            nn.init.zeros_(module.weight)
            nn.init.constant_(module.bias, head_bias)
        elif name.startswith("pre_logits"):
            lecun_normal_(module.weight)
            nn.init.zeros_(module.bias)
        elif jax_impl:
            nn.init.xavier_uniform_(module.weight)
            if module.bias is not None:
                if "mlp" in name:
                    nn.init.normal_(module.bias, std=1e-6)
                else:
                    nn.init.zeros_(module.bias)
        else:
            trunc_normal_(module.weight, std=0.02)
            if module.bias is not None:
                nn.init.zeros_(module.bias)
    elif jax_impl and isinstance(module, nn.Conv2d):
        lecun_normal_(module.weight)
        if module.bias is not None:
            nn.init.zeros_(module.bias)
    elif isinstance(module, (nn.LayerNorm, nn.GroupNorm, nn.BatchNorm2d)):
        nn.init.zeros_(module.bias)
        nn.init.ones_(module.weight)


def named_apply(fn: Callable, module: nn.Module, name: str = "", depth_first: bool = True, include_root: bool = False):
    if not depth_first and include_root:
        fn(module=module, name=name)
    for child_name, child in module.named_children():
        child_name = ".".join((name, child_name)) if name else child_name
        named_apply(fn=fn, module=child, name=child_name, depth_first=depth_first, include_root=True)
    if depth_first and include_root:
        fn(module=module, name=name)
    return module


class VisionTransformer(nn.Module):
    """timm-0.4.12 ``VisionTransformer`` surface (positional order is part of the contract:
    the reference calls ``super().__init__(img_size, patch_size, ..., weight_init)`` positionally)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, representation_size=None, distilled=False,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, embed_layer=PatchEmbed, norm_layer=None,
                 act_layer=None, weight_init=""):
        super().__init__()
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_tokens = 2 if distilled else 1
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        act_layer = act_layer or nn.GELU

        self.patch_embed = embed_layer(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches

        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.dist_token = nn.Parameter(torch.zeros(1, 1, embed_dim)) if distilled else None
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + self.num_tokens, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)

        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.Sequential(*[
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop_rate,
                  attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer, act_layer=act_layer)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim)

        if representation_size and not distilled:
            self.num_features = representation_size
            self.pre_logits = nn.Sequential(OrderedDict([
                ("fc", nn.Linear(embed_dim, representation_size)), ("act", nn.Tanh())]))
        else:
            self.pre_logits = nn.Identity()

        self.head = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        self.head_dist = None
        if distilled:
            self.head_dist = nn.Linear(self.embed_dim, self.num_classes) if num_classes > 0 else nn.Identity()

        self.init_weights(weight_init)

    # -- init -----------------------------------------------------------------------------
    def init_weights(self, mode=""):
        assert mode in ("jax", "jax_nlhb", "nlhb", "")
        head_bias = -math.log(self.num_classes) if "nlhb" in mode else 0.0
        trunc_normal_(self.pos_embed, std=0.02)
        if self.dist_token is not None:
            trunc_normal_(self.dist_token, std=0.02)
        if mode.startswith("jax"):
            named_apply(partial(_init_vit_weights, head_bias=head_bias, jax_impl=True), self)
        else:
            trunc_normal_(self.cls_token, std=0.02)
            self.apply(_init_vit_weights)

    def _init_weights(self, m):
        _init_vit_weights(m)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed", "cls_token", "dist_token"}

    def get_classifier(self):
        return self.head if self.dist_token is None else (self.head, self.head_dist)

    def reset_classifier(self, num_classes, global_pool=""):
        self.num_classes = num_classes
        self.head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        if self.num_tokens == 2:
            self.head_dist = nn.Linear(self.embed_dim, self.num_classes) if num_classes > 0 else nn.Identity()

    # -- forward --------------------------------------------------------------------------
    def embed(self, x):
        """patch-embed + cls (+dist) token + positional embedding."""
        return self.embed_tokens(self.patch_embed(x))

    def embed_tokens(self, x):
        """cat(cls [, dist], patches) + pos_embed -- handed to the first block unformed where its norm1 can fuse it
        (modules.Embedded)."""
        from . import modules
        if not getattr(self, "_defer_flags_set", False):
            # the models of this package own their block loops and materialise a deferred sum wherever something other than
            # the next block reads it (modules.value): their blocks may hand residual sums over un-added
            for blk in self.blocks:
                blk.defer_out = True
            self._defer_flags_set = True
        if x.dim() == 3 and not (self.training and self.pos_drop.p > 0) and len(self.blocks) > 0 and hasattr(self.blocks[0], "norm1"):
            tokens = self.cls_token[0] if self.dist_token is None else torch.cat((self.cls_token[0], self.dist_token[0]), dim=0)
            return modules.embed_tokens(x, tokens, self.pos_embed, self.blocks[0].norm1)
        cls_token = self.cls_token.expand(x.shape[0], -1, -1)
        if self.dist_token is None:
            x = torch.cat((cls_token, x), dim=1)
        else:
            x = torch.cat((cls_token, self.dist_token.expand(x.shape[0], -1, -1), x), dim=1)
        return self.pos_drop(x + self.pos_embed)

    def forward_features(self, x):
        x = self.embed(x)
        from . import modules
        x = modules.value(self.blocks(x))
        x = self.norm(x)
        if self.dist_token is None:
            return self.pre_logits(x[:, 0])
        return x[:, 0], x[:, 1]

    def classify(self, x):
        """final norm already applied; x = cls feature(s)."""
        if self.head_dist is not None:
            x, x_dist = self.head(x[0]), self.head_dist(x[1])
            if self.training and not torch.jit.is_scripting():
                return x, x_dist
            return (x + x_dist) / 2
        return self.head(x)

    def forward(self, x):
        return self.classify(self.forward_features(x))


# --------------------------------------------------------------------------- cfg + registry
def _cfg(url: str = "", **kwargs) -> Dict:
    cfg = {
        "url": url, "num_classes": 1000, "input_size": (3, 224, 224), "pool_size": None,
        "crop_pct": 0.9, "interpolation": "bicubic", "fixed_input_size": True,
        "mean": IMAGENET_DEFAULT_MEAN, "std": IMAGENET_DEFAULT_STD,
        "first_conv": "patch_embed.proj", "classifier": "head",
    }
    cfg.update(kwargs)
    return cfg


_DEIT = "https://dl.fbaipublicfiles.com/deit/"
default_cfgs: Dict[str, Dict] = {
    "deit_tiny_patch16_224": _cfg(url=_DEIT + "deit_tiny_patch16_224-a1311bcf.pth"),
    "deit_small_patch16_224": _cfg(url=_DEIT + "deit_small_patch16_224-cd65a155.pth"),
    "deit_base_patch16_224": _cfg(url=_DEIT + "deit_base_patch16_224-b5f2ef4d.pth"),
    "deit_tiny_distilled_patch16_224": _cfg(url=_DEIT + "deit_tiny_distilled_patch16_224-b40b3cf7.pth",
                                            classifier=("head", "head_dist")),
    "deit_small_distilled_patch16_224": _cfg(url=_DEIT + "deit_small_distilled_patch16_224-649709d9.pth",
                                             classifier=("head", "head_dist")),
    "deit_base_distilled_patch16_224": _cfg(url=_DEIT + "deit_base_distilled_patch16_224-df68dfff.pth",
                                            classifier=("head", "head_dist")),
}
