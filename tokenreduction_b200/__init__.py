"""tokenreduction_b200 — B200-native (sm_100a) token-reduction operators for DeiT, drop-in for the reduction
modules of JoakimHaurum/TokenReduction.  See DESIGN.md / INTEGRATION.md.

    from tokenreduction_b200 import create_model
    model = create_model("tome_small_patch16_224", num_classes=1000, args=Namespace(keep_rate=[0.7], reduction_loc=[3, 6, 9]))
"""
from .factory import create_model, is_model, list_models, register_model  # noqa: F401

__version__ = "0.1.0"
