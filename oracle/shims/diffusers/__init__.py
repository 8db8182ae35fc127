"""Stub package named ``diffusers`` - TEST INFRASTRUCTURE (see oracle/__init__.py).

Exposes exactly the names the reference imports (src/base_diffusion.py:4-8,
src/diffusion_classes.py:3-6, src/models.py:4-8, src/utils.py:4) so the unmodified
reference ``src/*.py`` can be imported in the dev container with
``sys.path = [<repo>/oracle/shims, /root/reference/src]``.  DDIMScheduler and
UNet2DModel are the oracle restatements; the remaining names are placeholders
(their models need hub checkpoints, which are unavailable offline)."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _root not in sys.path:
    sys.path.append(_root)

from oracle.ddim_scheduler import DDIMScheduler  # noqa: E402,F401
from oracle.unet2d import UNet2DModel  # noqa: E402,F401


class _Unavailable:
    def __init__(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__} needs hub weights; unavailable offline")

    @classmethod
    def from_pretrained(cls, *a, **k):
        raise NotImplementedError(f"{cls.__name__}.from_pretrained: no network")


class UNet2DConditionModel(_Unavailable):
    pass


class AutoencoderKL(_Unavailable):
    pass


class VQModel(_Unavailable):
    pass


class DiffusionPipeline(_Unavailable):
    pass


class StableDiffusionPipeline(_Unavailable):
    pass
