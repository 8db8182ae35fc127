"""Tensor <-> PIL conversion (drop-in for src/transforms.py).

tensor_to_pil: the scale/clamp/x255/truncate arithmetic of the reference
(``(x/2+0.5).clamp(0,1)`` -> ToPILImage's ``mul(255).byte()``, src/transforms.py:8-35) runs as one
CUDA kernel that writes HWC uint8; a single D->H copy follows."""
from typing import List, Union

import numpy as np
import torch
from PIL import Image

from b200edit import ops


def tensor_to_uint8(tensor_img: torch.Tensor) -> torch.Tensor:
    """(B,C,H,W) or (C,H,W) fp32 in [-1,1] -> (B,H,W,C) uint8 on the device."""
    if tensor_img.dim() == 3:
        tensor_img = tensor_img.unsqueeze(0)
    return ops.to_uint8(tensor_img)


def _to_pil(u8_hwc: np.ndarray) -> Image.Image:
    if u8_hwc.shape[-1] == 1:
        return Image.fromarray(u8_hwc[..., 0], mode="L")
    return Image.fromarray(u8_hwc)


def tensor_to_pil(tensor_img: torch.Tensor) -> Image.Image:
    if tensor_img.dim() == 2:          # integer mask: cast, no scaling (src/transforms.py:19-24)
        return Image.fromarray(tensor_img.detach().cpu().to(torch.uint8).numpy(), mode="L")
    if tensor_img.dim() == 4:
        assert tensor_img.shape[0] == 1
    elif tensor_img.dim() != 3:
        raise Exception("Input tensor has wrong shape")
    return _to_pil(tensor_to_uint8(tensor_img)[0].cpu().numpy())


def tensors_to_pils(tensor_imgs: List[torch.Tensor]) -> List[Image.Image]:
    return [tensor_to_pil(t) for t in tensor_imgs]


def batch_to_pils(batch: torch.Tensor) -> List[Image.Image]:
    """(B,C,H,W) -> B PIL images with ONE kernel and ONE device->host copy."""
    dev = tensor_to_uint8(batch)
    if dev.is_cuda:
        # pinned staging buffer (cached per shape): the copy runs at PCIe speed instead of through a pageable bounce buffer
        host = _pinned_u8(tuple(dev.shape))
        host.copy_(dev, non_blocking=True)
        torch.cuda.current_stream(dev.device).synchronize()
        u8 = host.numpy()
    else:
        u8 = dev.numpy()
    return [_to_pil(u8[i].copy()) for i in range(u8.shape[0])]      # copies: the staging buffer is reused by the next call


_PINNED = {}


def _pinned_u8(shape):
    buf = _PINNED.get(shape)
    if buf is None:
        if len(_PINNED) >= 4:
            _PINNED.clear()
        buf = _PINNED[shape] = torch.empty(shape, dtype=torch.uint8).pin_memory()
    return buf


def pil_to_tensor(pil_imgs: Union[Image.Image, List[Image.Image]]) -> torch.Tensor:
    """PIL -> (B,C,H,W) fp32 in [-1,1] on the host (x/255*2-1, src/transforms.py:48-65)."""
    def one(img):
        a = np.asarray(img)
        if a.ndim == 2:
            a = a[:, :, None]
        t = torch.from_numpy(np.array(a, copy=True)).permute(2, 0, 1).to(torch.float32).div(255)
        return (t * 2 - 1).unsqueeze(0)
    if isinstance(pil_imgs, Image.Image):
        return one(pil_imgs)
    if isinstance(pil_imgs, list):
        return torch.cat([one(i) for i in pil_imgs])
    raise Exception("Input need to be PIL.Image or list of PIL.Image")
