"""timm-style model factory with the reference's entrypoint names (models_act.py:8-51).

``create_model(name, pretrained=False, num_classes=..., drop_rate=..., drop_path_rate=..., drop_block_rate=None,
img_size=224, args=Namespace)`` has the call shape of train.py:322-331 / validate.py:88-94; like timm 0.4.12 it
drops None-valued kwargs before calling the entrypoint.  ``args`` carries ``keep_rate: List[float]``,
``reduction_loc: List[int]`` and the method flags (k_neighbors, cluster_iters, sinkhorn_eps, equal_weight,
dyvit_distill, distillation_type, viz_mode).  Pretrained DeiT weights are downloaded by the reference
(models_act.py:54-60); there is no network here, so ``pretrained=True`` raises.
"""
from __future__ import annotations

from functools import partial
from typing import Callable, Dict

import torch.nn as nn

from . import models as R
from .vit import VisionTransformer, default_cfgs

_SIZES = {"tiny": dict(embed_dim=192, num_heads=3), "small": dict(embed_dim=384, num_heads=6),
          "base": dict(embed_dim=768, num_heads=12)}

_METHODS = {
    "topk": R.TopKVisionTransformer, "evit": R.EfficientVisionTransformer, "tome": R.ToMeVisionTransformer,
    "dyvit": R.DynamicVisionTransformer, "dpcknn": R.DPCKNNVisionTransformer, "kmedoids": R.KMedoidsVisionTransformer,
    "sinkhorn": R.SinkhornVisionTransformer, "patchmerger": R.PatchMergerVisionTransformer,
    "ats": R.ATSVisionTransformer, "sit": R.SelfSlimmedVisionTransformer,
}

_model_entrypoints: Dict[str, Callable] = {}


def register_model(fn: Callable) -> Callable:
    _model_entrypoints[fn.__name__] = fn
    return fn


def list_models():
    return sorted(_model_entrypoints)


def is_model(name: str) -> bool:
    return name in _model_entrypoints


def create_model(model_name: str, pretrained: bool = False, checkpoint_path: str = "", scriptable=None,
                 exportable=None, no_jit=None, **kwargs):
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    if model_name not in _model_entrypoints:
        raise RuntimeError("Unknown model (%s)" % model_name)
    return _model_entrypoints[model_name](pretrained=pretrained, **kwargs)


def _distilled(kwargs) -> bool:
    args = kwargs.get("args")
    return hasattr(args, "distillation_type") and args.distillation_type != "none"


def _finish(model, size: str, distilled: bool, pretrained: bool):
    key = f"deit_{size}_distilled_patch16_224" if distilled else f"deit_{size}_patch16_224"
    model.default_cfg = default_cfgs[key]
    if pretrained:
        raise RuntimeError("pretrained DeiT weights are fetched from a URL by the reference (models_act.py:54-60); "
                           "no network in this environment — load a state_dict explicitly")
    return model


def _make_entry(method: str, size: str):
    cls = _METHODS[method]

    def entry(pretrained=False, **kwargs):
        distilled = _distilled(kwargs)
        extra = {}
        if method == "dyvit":     # models_act.py:284-294: reads args.dyvit_distill and forwards `distilled`
            extra = dict(dyvit_distillation=kwargs["args"].dyvit_distill, distilled=distilled)
        model = cls(patch_size=16, depth=12, mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                    **_SIZES[size], **extra, **kwargs)
        return _finish(model, size, distilled, pretrained)

    entry.__name__ = f"{method}_{size}_patch16_224"
    entry.__doc__ = f"{cls.__name__} at DeiT-{size} (models_act.py entrypoint of the same name)."
    return register_model(entry)


def _make_deit(size: str):
    def entry(pretrained=False, **kwargs):        # models_act.py:63-170: plain DeiT, `args` only read for distillation
        distilled = _distilled(kwargs)
        kwargs.pop("args", None)
        model = VisionTransformer(patch_size=16, depth=12, mlp_ratio=4, qkv_bias=True,
                                  norm_layer=partial(nn.LayerNorm, eps=1e-6), distilled=distilled, **_SIZES[size], **kwargs)
        return _finish(model, size, distilled, pretrained)

    entry.__name__ = f"deit_{size}_patch16_224_local"
    return register_model(entry)


def _make_out_of_scope(name: str, why: str):
    def entry(pretrained=False, **kwargs):
        raise NotImplementedError(f"{name}: {why}")

    entry.__name__ = name
    return register_model(entry)


for _size in _SIZES:
    _make_deit(_size)
    for _method in _METHODS:
        _make_entry(_method, _size)
    # the remaining reference entrypoints have no reduction operator to accelerate
    _make_out_of_scope(f"heuristic_{_size}_patch16_224", "static attention masks, no token is removed or merged "
                       "(models/heuristic.py:231-259) — use the reference implementation")
    _make_out_of_scope(f"deit_{_size}_patch16_224_local_viz", "feature-dump variant of plain DeiT (models/deit_viz.py)")
    _make_out_of_scope(f"dyvit_{_size}_patch16_224_teacher", "DynamicViT distillation teacher (training only)")

__all__ = ["create_model", "register_model", "list_models", "is_model"] + sorted(_model_entrypoints)
globals().update(_model_entrypoints)
