"""Grey-scale morphology operators (drop-in for src/Morphology.py): k x k dilation / erosion with a
learnable (zero-initialised) structuring element, zero 'same' padding, hard max or soft
log-sum-exp.  forward() is one CUDA kernel (inference only - no autograd)."""
import torch
import torch.nn as nn

from b200edit import ops


class Morphology(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=5, soft_max=True, beta=15, type=None):
        super().__init__()
        if type not in ("dilation2d", "erosion2d"):
            raise ValueError(f"unknown morphology type {type!r}")
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.soft_max, self.beta, self.type = soft_max, beta, type
        self.weight = nn.Parameter(torch.zeros(out_channels, in_channels, kernel_size, kernel_size))

    @torch.no_grad()
    def forward(self, x):
        return ops.morphology2d(x, self.weight.detach().to(x.device), self.type, self.soft_max, self.beta)


class Dilation2d(Morphology):
    def __init__(self, in_channels, out_channels, kernel_size=5, soft_max=True, beta=20):
        super().__init__(in_channels, out_channels, kernel_size, soft_max, beta, "dilation2d")


class Erosion2d(Morphology):
    def __init__(self, in_channels, out_channels, kernel_size=5, soft_max=True, beta=20):
        super().__init__(in_channels, out_channels, kernel_size, soft_max, beta, "erosion2d")


def fixed_padding(inputs, kernel_size, dilation):
    k_eff = kernel_size + (kernel_size - 1) * (dilation - 1)
    beg = (k_eff - 1) // 2
    return torch.nn.functional.pad(inputs, (beg, k_eff - 1 - beg, beg, k_eff - 1 - beg))
