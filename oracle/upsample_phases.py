"""Sub-pixel phase decomposition of "nearest-upsample x2, then 3x3 convolution (padding 1)" - TEST INFRASTRUCTURE.

diffusers' ``Upsample2D`` (the ``upsamplers.0`` of every UNet / decoder up block the reference runs through
``model.unet(...)`` / ``model.decode(...)``, src/diffusion_utils.py:27-31, src/diffusion_classes.py:45-70) is
``conv3x3(interpolate(x, scale_factor=2, mode="nearest"))``.  Output pixel (2i + a, 2j + b) of that composition only sees
the input rows {i + a - 1, i + a} and columns {j + b - 1, j + b}: the 3x3 taps collapse into 2x2 taps whose weights are
sums of the original ones.  The engine computes the four phases as four 2x2 convolutions over the LOW-resolution tensor
(csrc/conv_igemm.cu: ``conv_pack_weight_up2`` / ``ConvDesc::up2_phase``); this module restates the weight folding and the
phase assembly so that the identity is pinned on the CPU, independent of the CUDA code."""
from __future__ import annotations

import torch
import torch.nn.functional as F

# 3x3 rows (or columns) folded into tap t of phase a: FOLD[a][t]
FOLD = (((0,), (1, 2)), ((0, 1), (2,)))


def phase_weights(w: torch.Tensor, a: int, b: int) -> torch.Tensor:
    """w (Cout, Cin, 3, 3) -> the 2x2 weights (Cout, Cin, 2, 2) of phase (a, b); taps summed in the order the pack kernel uses."""
    out = torch.zeros(w.shape[0], w.shape[1], 2, 2, dtype=w.dtype)
    for th in range(2):
        for tw in range(2):
            acc = torch.zeros(w.shape[0], w.shape[1], dtype=w.dtype)
            for kh in FOLD[a][th]:
                for kw in FOLD[b][tw]:
                    acc = acc + w[:, :, kh, kw]
            out[:, :, th, tw] = acc
    return out


def upsample_conv_by_phases(x: torch.Tensor, w: torch.Tensor, bias=None) -> torch.Tensor:
    """x (N, Cin, H, W) -> (N, Cout, 2H, 2W): four 2x2 convolutions over x, tap (th, tw) of phase (a, b) reading the input at
    (i + a - 1 + th, j + b - 1 + tw) with zero padding, written to the output pixels (2i + a, 2j + b)."""
    N, _, H, W = x.shape
    out = torch.empty(N, w.shape[0], 2 * H, 2 * W, dtype=x.dtype)
    for a in range(2):
        for b in range(2):
            # pad so that a plain (valid) 2x2 correlation starts at input offset (a - 1, b - 1)
            xp = F.pad(x, (1 - b, b, 1 - a, a))
            out[:, :, a::2, b::2] = F.conv2d(xp, phase_weights(w, a, b), bias)
    return out


def upsample_conv_reference(x: torch.Tensor, w: torch.Tensor, bias=None) -> torch.Tensor:
    return F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, bias, padding=1)
