"""TEST INFRASTRUCTURE ONLY — a stand-in for the parts of timm==0.4.12 that the reference imports.

The reference (/root/reference, pinned to timm==0.4.12 in requirements.txt:8) cannot be imported in
this image because timm is absent and there is no network.  This module registers a synthetic ``timm``
package tree in ``sys.modules`` that re-exports the backbone in ``tokenreduction_b200.vit`` under the
names the reference imports (models/topk.py:8-11, models_act.py:5-6, models/deit_viz.py:11-13).  With
it, the UNMODIFIED reference modules import and run on CPU, which is how the oracle restatement
(oracle/ops.py, oracle/model.py) is validated and how tests/golden/ vectors are generated here.

Nothing under tokenreduction_b200/ imports this file.
"""
from __future__ import annotations

import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference_root() -> str:
    """TOKRED_REFERENCE_ROOT, else the read-only mount of the build container, else the git-ignored staging copy
    baseline/_ref that __graft_entry__.build() makes there (it travels to the GPU box; the mount does not)."""
    cands = [os.environ.get("TOKRED_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "models")) and os.path.isfile(os.path.join(c, "models_act.py")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_reference_root()

_model_entrypoints = {}


def register_model(fn):
    _model_entrypoints[fn.__name__] = fn
    return fn


def create_model(model_name, pretrained=False, checkpoint_path="", scriptable=None, exportable=None, no_jit=None, **kwargs):
    # timm 0.4.12 drops None-valued kwargs before calling the entrypoint (so drop_block_rate=None never
    # reaches the reference constructors, train.py:322-331).
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    if model_name not in _model_entrypoints:
        raise RuntimeError("Unknown model (%s)" % model_name)
    return _model_entrypoints[model_name](pretrained=pretrained, **kwargs)


def install() -> None:
    """Idempotently register the fake ``timm`` tree."""
    if "timm" in sys.modules and getattr(sys.modules["timm"], "__tokred_shim__", False):
        return
    from tokenreduction_b200 import vit

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    timm = mod("timm", __tokred_shim__=True, __version__="0.4.12")
    timm.__path__ = []  # mark as package
    data = mod("timm.data", IMAGENET_DEFAULT_MEAN=vit.IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD=vit.IMAGENET_DEFAULT_STD)
    data.__path__ = []
    mod("timm.data.constants", IMAGENET_DEFAULT_MEAN=vit.IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD=vit.IMAGENET_DEFAULT_STD)
    models = mod("timm.models", create_model=create_model)
    models.__path__ = []
    layers = mod("timm.models.layers", PatchEmbed=vit.PatchEmbed, Mlp=vit.Mlp, DropPath=vit.DropPath,
                 trunc_normal_=vit.trunc_normal_, lecun_normal_=vit.lecun_normal_)
    registry = mod("timm.models.registry", register_model=register_model)

    def _unsupported(*a, **k):
        raise NotImplementedError("timm shim: helper is import-only")

    helpers = mod("timm.models.helpers", build_model_with_cfg=_unsupported, named_apply=vit.named_apply,
                  adapt_input_conv=_unsupported)
    vt = mod("timm.models.vision_transformer", VisionTransformer=vit.VisionTransformer, _cfg=vit._cfg,
             default_cfgs=vit.default_cfgs, Block=vit.Block, Attention=vit.Attention)
    timm.data, timm.models = data, models
    models.layers, models.registry, models.helpers, models.vision_transformer = layers, registry, helpers, vt


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models")) and os.path.isfile(os.path.join(REFERENCE_ROOT, "models_act.py"))


def import_reference():
    """Import the unmodified reference (``models.*`` namespace package + ``models_act``). Container-only."""
    if not reference_available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    models_act = importlib.import_module("models_act")
    return models_act
