"""Oracle restatement (numpy) of the mask path: src/mask_creator.py + src/Morphology.py.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PINNED against the reference run on
its own fixture parsing maps (tests/golden/mask_*.npz, made by make_golden.py).

``MaskCreator.create_mask`` (src/mask_creator.py:22-55):
  per class  m_c = (seg == c).float()                                   (:34)
  optional   m_c = Dilation2d(1,1,7,soft_max=False)(m_c)                (:16,:36-40)
  m = sum_c m_c                                                         (:26)
  r = torchvision Resize((d,d))(m[None])  -> bilinear, antialias=True   (:20,:51)
  r[r<1] = 0 ; r[r>1] = 1   (exactly-1 stays 1)                         (:52-53)
  mask = stack 3x -> (1,3,d,d)                                          (:43-48)

The antialiased resize is ATen's separable triangle filter
(aten/src/ATen/native/cpu/UpSampleKernel.cpp, ``_upsample_bilinear2d_aa``), which the
reference reaches through torchvision 0.26's ``Resize`` default.  It is not part of
/root/reference; its arithmetic was established by probing torch 2.11 CPU in the dev
container (tests/golden/make_golden.py records torch's own output as the golden):
  * horizontal (W) pass first, then vertical (H), fp32 intermediates;
  * per output: weights w_j = tri((j + xmin - center + 0.5) * invscale) / sum, fp32;
  * accumulation is strictly sequential: acc = x_0*w_0, then for j = 1..L-1 the first
    4*floor((L-1)/4) terms are added as round(x_j*w_j) (separate multiply and add) and
    the remaining (L-1) mod 4 terms as fused multiply-adds - the shape of the compiled
    loop (4-wide vectorised multiply with an in-order reduction + scalar FMA tail).
For {0,1,2..}-valued masks and power-of-two ratios every interior value is dyadic and
exact under any order; the order only matters where the filter support is clipped by
the image border (weights such as 3/7).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def aa_weights(in_size: int, out_size: int):
    """[(xmin, [w_j...])] per output index.  Mirrors the float/double promotions of ATen's
    ``_compute_indices_min_size_weights_aa`` with scalar_t = float: literals such as 0.5 and
    1.0 are doubles in the C++ source, stored results are floats."""
    F64 = np.float64
    scale = F32(F32(in_size) / F32(out_size))
    support = F32(F64(1.0) * F64(scale)) if scale >= 1 else F32(1.0)
    invscale = F32(F64(1.0) / F64(scale)) if scale >= 1 else F32(1.0)
    max_taps = int(np.ceil(support)) * 2 + 1
    table = []
    for i in range(out_size):
        center = F32(F64(scale) * (F64(i) + F64(0.5)))
        xmin = max(int(F64(F32(center - support)) + F64(0.5)), 0)
        xsize = min(int(F64(F32(center + support)) + F64(0.5)), in_size) - xmin
        xsize = min(max(xsize, 0), max_taps)
        ws, total = [], F32(0)
        for j in range(xsize):
            d = F32(F32(j + xmin) - center)
            x = abs(F32((F64(d) + F64(0.5)) * F64(invscale)))
            w = F32(F64(1.0) - F64(x)) if x < 1 else F32(0)
            ws.append(w)
            total = F32(total + w)
        if total != 0:
            ws = [F32(w / total) for w in ws]
        table.append((xmin, ws))
    return table


def _fma(a, b, c):
    # a*b is exact in float64 (24+24 significant bits); one rounding to fp32 follows.
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(F32)


def resize_last_dim_aa(x: np.ndarray, out_size: int) -> np.ndarray:
    """x: (rows, n) fp32 -> (rows, out_size) with the ATen accumulation order."""
    table = aa_weights(x.shape[1], out_size)
    out = np.zeros((x.shape[0], out_size), F32)
    for i, (xmin, ws) in enumerate(table):
        acc = (x[:, xmin] * ws[0]).astype(F32)
        n_plain = 4 * ((len(ws) - 1) // 4)
        for j in range(1, len(ws)):
            if j - 1 < n_plain:
                acc = (acc + (x[:, xmin + j] * ws[j]).astype(F32)).astype(F32)
            else:
                acc = _fma(x[:, xmin + j], ws[j], acc)
        out[:, i] = acc
    return out


def resize_bilinear_aa(x: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """(H,W) fp32 -> (out_h,out_w): W pass, then H pass."""
    t = resize_last_dim_aa(np.ascontiguousarray(x, F32), out_w)
    return np.ascontiguousarray(resize_last_dim_aa(np.ascontiguousarray(t.T), out_h).T)


def dilate_hard(x: np.ndarray, k: int = 7, weight: np.ndarray | None = None) -> np.ndarray:
    """Grey dilation with zero padding k//2 and hard max (src/Morphology.py:47-84,105-111);
    ``weight`` (k,k) is the structuring element (zero-initialised in the reference)."""
    h, w = x.shape
    p = k // 2
    pad = np.zeros((h + k - 1, w + k - 1), F32)
    pad[p:p + h, p:p + w] = x
    out = np.full((h, w), -np.inf, F32)
    for dy in range(k):
        for dx in range(k):
            wv = F32(0) if weight is None else F32(weight[dy, dx])
            out = np.maximum(out, (pad[dy:dy + h, dx:dx + w] + wv).astype(F32))
    return out


def erode_hard(x: np.ndarray, k: int = 7, weight: np.ndarray | None = None) -> np.ndarray:
    """Grey erosion: -max(weight - x) over the zero-padded window (src/Morphology.py:64,80)."""
    h, w = x.shape
    p = k // 2
    pad = np.zeros((h + k - 1, w + k - 1), F32)
    pad[p:p + h, p:p + w] = x
    out = np.full((h, w), -np.inf, F32)
    for dy in range(k):
        for dx in range(k):
            wv = F32(0) if weight is None else F32(weight[dy, dx])
            out = np.maximum(out, (wv - pad[dy:dy + h, dx:dx + w]).astype(F32))
    return (-out).astype(F32)


def create_mask(seg: np.ndarray, classes, dilate: bool, size, channels: int = 3) -> np.ndarray:
    """seg (H,W) integer class map -> (1,channels,d_h,d_w) fp32 in {0,1}."""
    total = np.zeros(seg.shape, F32)
    for c in classes:
        m = (seg == c).astype(F32)
        if dilate:
            m = dilate_hard(m, 7)
        total = (total + m).astype(F32)
    r = resize_bilinear_aa(total, size[0], size[1])
    out = (r >= 1).astype(F32)
    return np.broadcast_to(out, (1, channels) + out.shape).copy()
