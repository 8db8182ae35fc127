"""Oracle restatement of ``diffusers.UNet2DModel`` (the DDPM / LDM noise predictor).

TEST INFRASTRUCTURE.  PARITY UNPINNED (see oracle/__init__.py): the UNet lives in
the third-party ``diffusers`` package, which is neither vendored under
/root/reference nor installed here.  Restated from the published layout of
``google/ddpm-celebahq-256`` (config.json of the hub model the reference loads at
src/models.py:21): positional sin/cos timestep embedding -> 2-layer MLP; a
ResNet/attention encoder-decoder with skip concatenation; GroupNorm(32, eps 1e-6)
+ SiLU everywhere; single-head self-attention at 16x16 and in the mid block.

Reference call sites: ``model.unet(latent, t)["sample"]`` src/diffusion_utils.py:72;
``unet.config.in_channels / sample_size`` src/utils.py:68-70 and the deprecated
``unet.in_channels / sample_size`` src/ddpm_inversion.py:39-41, src/base_diffusion.py:38.

Parameter names follow diffusers' state_dict so a real checkpoint could be loaded.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

DDPM256_CONFIG = dict(
    sample_size=256, in_channels=3, out_channels=3,
    block_out_channels=(128, 128, 256, 256, 512, 512), layers_per_block=2,
    down_block_types=("DownBlock2D",) * 4 + ("AttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "AttnUpBlock2D") + ("UpBlock2D",) * 4,
    norm_num_groups=32, norm_eps=1e-6, attention_head_dim=None,
    flip_sin_to_cos=False, freq_shift=1,
)


# CompVis/ldm-celebahq-256 `unet` (UNet2DModel on the 64x64x3 VQ latent); recalled layout, unverifiable offline
LDM_CELEBAHQ_CONFIG = dict(
    sample_size=64, in_channels=3, out_channels=3,
    block_out_channels=(224, 448, 672, 896), layers_per_block=2,
    down_block_types=("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"),
    up_block_types=("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
    norm_num_groups=32, norm_eps=1e-6, attention_head_dim=32,
    flip_sin_to_cos=True, freq_shift=0, downsample_padding=1,
)


def timestep_embedding(timesteps, dim, flip_sin_to_cos=False, freq_shift=1, max_period=10000):
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32,
                                                    device=timesteps.device)
    exponent = exponent / (half - freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cout)
        self.linear_2 = nn.Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_ch, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_ch, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    """Self-attention block with GroupNorm and residual (heads = C // head_dim, 1 if None)."""

    def __init__(self, ch, head_dim, groups, eps):
        super().__init__()
        self.heads = 1 if head_dim is None else ch // head_dim
        self.group_norm = nn.GroupNorm(groups, ch, eps=eps)
        self.to_q = nn.Linear(ch, ch)
        self.to_k = nn.Linear(ch, ch)
        self.to_v = nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch)])

    def forward(self, x):
        b, c, hh, ww = x.shape
        h = self.group_norm(x).view(b, c, hh * ww).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        d = c // self.heads

        def split(t):
            return t.view(b, hh * ww, self.heads, d).transpose(1, 2)

        q, k, v = split(q), split(k), split(v)
        p = torch.softmax((q @ k.transpose(-1, -2)) * (d ** -0.5), dim=-1)
        o = (p @ v).transpose(1, 2).reshape(b, hh * ww, c)
        o = self.to_out[0](o)
        return o.transpose(1, 2).reshape(b, c, hh, ww) + x


class Downsample2D(nn.Module):
    """diffusers Downsample2D(use_conv=True, padding=p): p = 0 pads (0,1,0,1) first (DDPM-256 config), p = 1 is a
    plain padding-1 stride-2 convolution (UNet2DModel default, LDM config)."""

    def __init__(self, ch, padding=0):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1))
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb_ch, layers, groups, eps, attn, head_dim, add_down, down_pad=0):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, temb_ch, groups, eps) for i in range(layers)])
        self.attentions = nn.ModuleList(
            [Attention(cout, head_dim, groups, eps) for _ in range(layers)]) if attn else None
        self.downsamplers = nn.ModuleList([Downsample2D(cout, down_pad)]) if add_down else None

    def forward(self, x, temb):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class UpBlock(nn.Module):
    def __init__(self, cin, cout, prev, temb_ch, layers, groups, eps, attn, head_dim, add_up):
        super().__init__()
        rs = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            rin = prev if i == 0 else cout
            rs.append(ResnetBlock2D(rin + skip, cout, temb_ch, groups, eps))
        self.resnets = nn.ModuleList(rs)
        self.attentions = nn.ModuleList(
            [Attention(cout, head_dim, groups, eps) for _ in range(layers)]) if attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, temb):
        for i, r in enumerate(self.resnets):
            x = r(torch.cat([x, skips.pop()], dim=1), temb)
            if self.attentions is not None:
                x = self.attentions[i](x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class MidBlock(nn.Module):
    def __init__(self, ch, temb_ch, groups, eps, head_dim):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb_ch, groups, eps) for _ in range(2)])
        self.attentions = nn.ModuleList([Attention(ch, head_dim, groups, eps)])

    def forward(self, x, temb):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x)
        return self.resnets[1](x, temb)


class UNet2DOutput(dict):
    @property
    def sample(self):
        return self["sample"]


class UNet2DModel(nn.Module):
    def __init__(self, sample_size=256, in_channels=3, out_channels=3,
                 block_out_channels=(128, 128, 256, 256, 512, 512), layers_per_block=2,
                 down_block_types=None, up_block_types=None, norm_num_groups=32,
                 norm_eps=1e-6, attention_head_dim=None, flip_sin_to_cos=False, freq_shift=1,
                 downsample_padding=0):
        super().__init__()
        n = len(block_out_channels)
        down_block_types = tuple(down_block_types or ("DownBlock2D",) * n)
        up_block_types = tuple(up_block_types or ("UpBlock2D",) * n)
        self.config = SimpleNamespace(
            sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
            block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
            down_block_types=down_block_types, up_block_types=up_block_types,
            norm_num_groups=norm_num_groups, norm_eps=norm_eps,
            attention_head_dim=attention_head_dim, flip_sin_to_cos=flip_sin_to_cos,
            freq_shift=freq_shift, downsample_padding=downsample_padding)
        # deprecated direct attributes the reference still reads
        self.in_channels = in_channels
        self.sample_size = sample_size
        boc = list(block_out_channels)
        temb_ch = boc[0] * 4
        g, eps, hd = norm_num_groups, norm_eps, attention_head_dim
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb_ch)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(DownBlock(in_ch, out_ch, temb_ch, layers_per_block, g, eps,
                                              t.startswith("Attn"), hd, i != n - 1, downsample_padding))
        self.mid_block = MidBlock(boc[-1], temb_ch, g, eps, hd)
        self.up_blocks = nn.ModuleList()
        rev = boc[::-1]
        out_ch = rev[0]
        for i, t in enumerate(up_block_types):
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, n - 1)]
            self.up_blocks.append(UpBlock(in_ch, out_ch, prev, temb_ch, layers_per_block + 1, g, eps,
                                          t.startswith("Attn"), hd, i != n - 1))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=eps)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    @property
    def device(self):
        return next(self.parameters()).device

    def forward(self, sample, timestep, **_):
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.long, device=sample.device)
        timestep = timestep.to(sample.device)
        if timestep.dim() == 0:
            timestep = timestep[None]
        timestep = timestep * torch.ones(sample.shape[0], dtype=timestep.dtype, device=sample.device)
        cfg = self.config
        temb = timestep_embedding(timestep, cfg.block_out_channels[0], cfg.flip_sin_to_cos,
                                  cfg.freq_shift).to(sample.dtype)
        temb = self.time_embedding(temb)
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb)
            skips += outs
        x = self.mid_block(x, temb)
        for blk in self.up_blocks:
            x = blk(x, skips, temb)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return UNet2DOutput(sample=x)


class ToyEpsModel(nn.Module):
    """A tiny, cheap, deterministic stand-in for a UNet (two 3x3 convs + timestep
    modulation) with the same call contract; used where a test or a golden vector
    needs *some* noise predictor but the UNet itself is not under test."""

    def __init__(self, in_channels=3, sample_size=32, hidden=8, seed=0):
        super().__init__()
        self.config = SimpleNamespace(in_channels=in_channels, sample_size=sample_size)
        self.in_channels = in_channels
        self.sample_size = sample_size
        g = torch.Generator().manual_seed(seed)
        self.w1 = nn.Parameter(torch.randn(hidden, in_channels, 3, 3, generator=g) * 0.2)
        self.w2 = nn.Parameter(torch.randn(in_channels, hidden, 3, 3, generator=g) * 0.2)

    def forward(self, sample, timestep, **_):
        t = torch.as_tensor(timestep, dtype=torch.float32, device=sample.device) / 1000.0
        h = torch.tanh(F.conv2d(sample, self.w1, padding=1) * (1.0 + t))
        return UNet2DOutput(sample=F.conv2d(h, self.w2, padding=1) + 0.1 * sample)
