"""b200edit - host-side Python layer over libb200edit.so (hand-written sm_100a CUDA).

PyTorch is used only as the tensor container / stream provider.  There is no CPU, Triton or
cuDNN fallback: every operator raises if the CUDA library is missing or the device is not a B200.
"""
from . import _C  # noqa: F401
from ._C import B2EError, launch_count, lib  # noqa: F401
