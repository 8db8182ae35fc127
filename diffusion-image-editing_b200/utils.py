"""Helpers of the editing path (drop-in for src/utils.py)."""
from typing import List, Optional

import torch
from PIL import Image
from tqdm import tqdm

from b200edit import ops
from transforms import batch_to_pils, tensor_to_pil


def to_pil_and_decode_batch_of_tensors(model, tensor: torch.Tensor) -> List[Image.Image]:
    """Decode a (T,C,H,W) stack and convert it to PIL.  The reference decodes and copies one image
    at a time (src/utils.py:11-14); here the whole stack is decoded and converted in one pass."""
    if tensor.dim() == 3:
        tensor = tensor.unsqueeze(0)
    return batch_to_pils(model.decode(tensor))


def process_lists_of_tensors(model, tensors: List[torch.Tensor]) -> List[Image.Image]:
    stacked = torch.stack(tensors, dim=0)
    if stacked.dim() == 5 and stacked.shape[1] == 1:      # list of (1,C,H,W), the reference's case
        return to_pil_and_decode_batch_of_tensors(model, stacked[:, 0])
    if stacked.dim() == 5:                                # batched extension: T lists of B images
        T, B = stacked.shape[:2]
        flat = to_pil_and_decode_batch_of_tensors(model, stacked.reshape(T * B, *stacked.shape[2:]))
        return [flat[i * B:(i + 1) * B] for i in range(T)]
    return to_pil_and_decode_batch_of_tensors(model, stacked)


def apply_mask(mask: torch.Tensor, zo: torch.Tensor, zv: torch.Tensor) -> torch.Tensor:
    """mask*zv + (1-mask)*zo: resynthesise the noise inside the mask (src/utils.py:23-28)."""
    return ops.apply_mask(mask, zo, zv)


def get_device(verbose: bool = False) -> torch.device:
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    if verbose:
        print(f"Using {device} as backend")
    return device


def _unet_device(unet) -> torch.device:
    d = getattr(unet, "device", None)
    return torch.device(d) if d is not None else get_device()


def generate_random_samples(num_samples: Optional[int], unet,
                            generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Standard-normal samples drawn on the HOST generator, then copied to the model's device - the
    reference's RNG stream (src/utils.py:58-74), minus its hard-coded "cuda"."""
    n = 1 if num_samples is None else num_samples
    cfg = unet.config
    shape = (n, cfg.in_channels, cfg.sample_size, cfg.sample_size)
    return torch.randn(shape, generator=generator).to(_unet_device(unet))


def initialize_random_samples(model, num_inference_steps: int, eta: float, generator: torch.Generator):
    xt = generate_random_samples(1, model.unet, generator=generator)
    zs = generate_random_samples(num_inference_steps, model.unet, generator=generator) if eta > 0 else None
    return xt, zs


def create_progress_bar(steps, show_progbar: bool):
    it = enumerate(steps)
    return tqdm(it, total=len(steps)) if show_progbar else it


def set_seed(seed: Optional[int]) -> torch.Generator:
    if seed is None:
        seed = int(torch.randint(int(1e6), (1,)))
    return torch.manual_seed(seed)
