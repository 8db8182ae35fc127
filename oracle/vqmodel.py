"""Oracle restatement of the decode and encode paths of ``diffusers.VQModel`` / ``AutoencoderKL`` (CompVis/ldm-celebahq-256 ``vqvae``).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: diffusers is not vendored / installed and the
reference holds no golden vectors for it; this file restates the published layout (recalled from diffusers 0.2x,
``models/vq_model.py``, ``models/vae.py``) and is the normative spec for the native decoder:

    LDM.decode (src/diffusion_classes.py:62-70): vqvae.decode(latent.float()).sample
    VQModel.decode(h): quant = VectorQuantizer(h) (nearest code, straight-through gradient)
                       -> post_quant_conv (1x1) -> Decoder -> sample
    Decoder: conv_in 3x3 (latent -> C_top) -> UNetMidBlock2D (resnet, single-head attention, resnet; no time
             embedding) -> UpDecoderBlock2D per level, top-down, layers_per_block + 1 resnets each, nearest x2
             upsample + 3x3 conv between levels -> GroupNorm -> SiLU -> conv_out 3x3.

State-dict names follow diffusers (``decoder.up_blocks.0.resnets.1.conv_shortcut.weight``, ``quantize.embedding.weight``,
``post_quant_conv.weight`` ...), so the native engine and this module load the same dictionary.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet2d import Attention, Downsample2D, Upsample2D

# CompVis/ldm-celebahq-256 vqvae: 64x64x3 latent -> 256x256x3 image (x4), 8192 codes of dimension 3
LDM_VQ_CONFIG = dict(latent_channels=3, out_channels=3, block_out_channels=(128, 256, 512), layers_per_block=2,
                     norm_num_groups=32, norm_eps=1e-6, num_vq_embeddings=8192, sample_size=64)


class ResnetNoTemb(nn.Module):
    def __init__(self, cin, cout, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return h + (x if self.conv_shortcut is None else self.conv_shortcut(x))


class MidBlock(nn.Module):
    def __init__(self, ch, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetNoTemb(ch, ch, groups, eps) for _ in range(2)])
        self.attentions = nn.ModuleList([Attention(ch, None, groups, eps)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class UpDecoderBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, eps, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetNoTemb(cin if i == 0 else cout, cout, groups, eps) for i in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return x if self.upsamplers is None else self.upsamplers[0](x)


class Decoder(nn.Module):
    def __init__(self, latent_channels, out_channels, block_out_channels, layers_per_block, groups, eps):
        super().__init__()
        rev = list(block_out_channels)[::-1]
        self.conv_in = nn.Conv2d(latent_channels, rev[0], 3, padding=1)
        self.mid_block = MidBlock(rev[0], groups, eps)
        blocks, prev = [], rev[0]
        for i, ch in enumerate(rev):
            blocks.append(UpDecoderBlock(prev, ch, layers_per_block + 1, groups, eps, i != len(rev) - 1))
            prev = ch
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(groups, rev[-1], eps=eps)
        self.conv_out = nn.Conv2d(rev[-1], out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class DownEncoderBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, eps, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetNoTemb(cin if i == 0 else cout, cout, groups, eps) for i in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, padding=0)]) if add_down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return x if self.downsamplers is None else self.downsamplers[0](x)


class Encoder(nn.Module):
    """diffusers ``Encoder`` (models/vae.py): conv_in 3x3 -> DownEncoderBlock2D per level (layers_per_block resnets,
    pad (0,1,0,1) + 3x3 stride-2 convolution between levels) -> UNetMidBlock2D (resnet, single-head attention, resnet)
    -> GroupNorm -> SiLU -> conv_out 3x3 (2 x latent channels when double_z)."""

    def __init__(self, in_channels, latent_channels, block_out_channels, layers_per_block, groups, eps, double_z):
        super().__init__()
        chs = list(block_out_channels)
        self.conv_in = nn.Conv2d(in_channels, chs[0], 3, padding=1)
        blocks, prev = [], chs[0]
        for i, ch in enumerate(chs):
            blocks.append(DownEncoderBlock(prev, ch, layers_per_block, groups, eps, i != len(chs) - 1))
            prev = ch
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = MidBlock(chs[-1], groups, eps)
        self.conv_norm_out = nn.GroupNorm(groups, chs[-1], eps=eps)
        self.conv_out = nn.Conv2d(chs[-1], 2 * latent_channels if double_z else latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(self.mid_block(x))))


class DiagonalGaussianDistribution:
    """diffusers ``DiagonalGaussianDistribution``: moments = [mean | logvar] along channels."""

    def __init__(self, moments):
        self.mean, self.logvar = torch.chunk(moments, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device if generator is None else generator.device)
        return self.mean + self.std * noise.to(self.mean.device)


class VectorQuantizer(nn.Module):
    """Nearest codebook entry (squared Euclidean distance, first index on ties), straight-through gradient."""

    def __init__(self, n_e, dim):
        super().__init__()
        self.embedding = nn.Embedding(n_e, dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)

    def indices(self, z):
        zf = z.permute(0, 2, 3, 1).reshape(-1, z.shape[1])
        e = self.embedding.weight
        d = None
        for c in range(zf.shape[1]):      # ((d0^2 + d1^2) + d2^2): the order the native kernel uses
            t = (zf[:, None, c] - e[None, :, c]) ** 2
            d = t if d is None else d + t
        return torch.argmin(d, dim=1)

    def forward(self, z):
        idx = self.indices(z)
        zq = self.embedding(idx).view(z.shape[0], z.shape[2], z.shape[3], z.shape[1]).permute(0, 3, 1, 2)
        return z + (zq - z).detach()


class VQModel(nn.Module):
    def __init__(self, latent_channels=3, out_channels=3, block_out_channels=(128, 256, 512), layers_per_block=2,
                 norm_num_groups=32, norm_eps=1e-6, num_vq_embeddings=8192, sample_size=64):
        super().__init__()
        self.config = SimpleNamespace(latent_channels=latent_channels, out_channels=out_channels,
                                      block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                                      norm_num_groups=norm_num_groups, norm_eps=norm_eps,
                                      num_vq_embeddings=num_vq_embeddings, sample_size=sample_size)
        # num_vq_embeddings == 0: AutoencoderKL decode path (post_quant_conv -> decoder, no quantiser)
        self.quantize = VectorQuantizer(num_vq_embeddings, latent_channels) if num_vq_embeddings > 0 else None
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups,
                               norm_eps)
        # encode path (LDM.encode / SD.encode, src/diffusion_classes.py:27-30, 55-60).  Built AFTER the decoder so the
        # decoder's random-init stream (and every decode golden) is unchanged.
        self.double_z = num_vq_embeddings == 0
        q = 2 * latent_channels if self.double_z else latent_channels
        self.encoder = Encoder(out_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups,
                               norm_eps, self.double_z)
        self.quant_conv = nn.Conv2d(q, q, 1)

    def encode(self, x):
        """VQModel.encode(x).latents = quant_conv(encoder(x)) (quantisation happens in decode);
        AutoencoderKL.encode(x).latent_dist = DiagonalGaussianDistribution(quant_conv(encoder(x)))."""
        h = self.quant_conv(self.encoder(x))
        if self.double_z:
            return SimpleNamespace(latent_dist=DiagonalGaussianDistribution(h))
        return SimpleNamespace(latents=h)

    def decode(self, h, force_not_quantize=False):
        quant = h if (force_not_quantize or self.quantize is None) else self.quantize(h)
        return SimpleNamespace(sample=self.decoder(self.post_quant_conv(quant)))


# CompVis/stable-diffusion-v1-x vae decode path (no quantiser): 64x64x4 -> 512x512x3
SD_VAE_CONFIG = dict(latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                     norm_num_groups=32, norm_eps=1e-6, num_vq_embeddings=0, sample_size=64)
