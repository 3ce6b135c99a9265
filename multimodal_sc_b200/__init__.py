"""multimodal_sc_b200 — B200-native (sm_100a) implementation of the SemCom-PPO hot path.

Upstream (/root/reference/README.md:1-2) is a title only; the path is SPEC v0
(`multimodal_sc_b200.spec`).  Everything numeric runs in the
hand-written CUDA library `csrc/` -> `libsemcom_sm100.so`, reached through the C ABI
declared in `include/semcom_sm100.h`.  There is no CPU or eager fallback: calling an
op without the library (or without an sm_100 GPU) raises.
"""
from .spec import SpecV0  # noqa: F401

__all__ = ["SpecV0"]
__version__ = "0.1.0"
