"""playableenvironments_b200 — B200-native volumetric renderer behind the PlayableEnvironments module API.

Only the per-frame NeRF render path is implemented: the modules under ``model/`` and
``utils/`` mirror the reference's names, constructor arguments, ``forward`` signatures, ``state_dict`` keys
and config-string registry, and evaluate through ``libpe_b200.so`` (hand-written sm_100a kernels).
"""
from . import registry  # noqa: F401

__all__ = ["registry"]
