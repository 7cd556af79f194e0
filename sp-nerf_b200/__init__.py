"""spnerf_b200 — B200-native ray-rendering hot path of SP-NeRF.

Host-side mirror of the reference call surface (``modules.rendering.render_rays``,
``models.spnerf.{SPNeRF, inference}``, ``models.load_model``, ``modules.metrics`` losses) over the
C ABI in ``include/spnerf_b200.h`` (``lib/libspnerf_sm100a.so``, hand-written sm_100a CUDA).
There is no CPU path: every compute entry point raises if the library or a CUDA device is missing.
"""
from . import _cabi  # noqa: F401

__all__ = ["_cabi"]
