"""Oracle restatement of ``diffusers.UNet2DConditionModel`` in the Stable Diffusion 1.x layout.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: diffusers is not vendored / installed and the
reference holds no golden vectors for it; this file restates the published layout (recalled from diffusers 0.2x:
``models/unet_2d_condition.py``, ``unet_2d_blocks.py``, ``attention.py``) and is the normative spec of the native
conditional UNet.  The reference calls it as ``unet(cat([latent] * 2), t, encoder_hidden_states=text_emb)["sample"]``
(src/diffusion_utils.py:62-66).

    conv_in -> [CrossAttnDownBlock2D x3, DownBlock2D] -> UNetMidBlock2DCrossAttn -> [UpBlock2D, CrossAttnUpBlock2D x3]
    -> GroupNorm -> SiLU -> conv_out
    Transformer2DModel: GroupNorm(eps 1e-6) -> 1x1 conv proj_in -> BasicTransformerBlock -> 1x1 conv proj_out -> + x
    BasicTransformerBlock: x += attn1(LN(x)) ; x += attn2(LN(x), context) ; x += GEGLU-FF(LN(x))
    attention: 8 heads, head_dim = C / 8, to_q / to_k / to_v without bias, to_out.0 with bias

State-dict names follow diffusers (``down_blocks.0.attentions.1.transformer_blocks.0.attn2.to_k.weight`` ...).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet2d import Downsample2D, ResnetBlock2D, TimestepEmbedding, UNet2DOutput, Upsample2D, timestep_embedding

SD15_CONFIG = dict(sample_size=64, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280),
                   layers_per_block=2, cross_attention_dim=768, attention_head_dim=8, norm_num_groups=32, norm_eps=1e-5)


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim, heads):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(query_dim, query_dim, bias=False)
        self.to_k = nn.Linear(context_dim, query_dim, bias=False)
        self.to_v = nn.Linear(context_dim, query_dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(query_dim, query_dim)])

    def forward(self, x, context=None):
        ctx = x if context is None else context
        b, t, c = x.shape
        d = c // self.heads

        def split(u):
            return u.view(b, -1, self.heads, d).transpose(1, 2)

        q, k, v = split(self.to_q(x)), split(self.to_k(ctx)), split(self.to_v(ctx))
        p = torch.softmax((q @ k.transpose(-1, -2)) * (d ** -0.5), dim=-1)
        return self.to_out[0]((p @ v).transpose(1, 2).reshape(b, t, c))


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, context_dim):
        super().__init__()
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.attn1 = CrossAttention(dim, dim, heads)
        self.attn2 = CrossAttention(dim, context_dim, heads)
        self.ff = FeedForward(dim)

    def forward(self, x, context):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), context)
        return x + self.ff(self.norm3(x))


class Transformer2DModel(nn.Module):
    def __init__(self, ch, heads, context_dim, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.proj_in = nn.Conv2d(ch, ch, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(ch, heads, context_dim)])
        self.proj_out = nn.Conv2d(ch, ch, 1)

    def forward(self, x, context):
        b, c, hh, ww = x.shape
        h = self.proj_in(self.norm(x)).permute(0, 2, 3, 1).reshape(b, hh * ww, c)
        h = self.transformer_blocks[0](h, context)
        h = h.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
        return self.proj_out(h) + x


class CondDownBlock(nn.Module):
    def __init__(self, cin, cout, temb_ch, layers, groups, eps, attn, heads, ctx, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb_ch, groups, eps) for i in range(layers)])
        self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, ctx, groups) for _ in range(layers)]) if attn else None
        self.downsamplers = nn.ModuleList([Downsample2D(cout, 1)]) if add_down else None

    def forward(self, x, temb, context):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, context)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class CondUpBlock(nn.Module):
    def __init__(self, cin, cout, prev, temb_ch, layers, groups, eps, attn, heads, ctx, add_up):
        super().__init__()
        rs = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            rin = prev if i == 0 else cout
            rs.append(ResnetBlock2D(rin + skip, cout, temb_ch, groups, eps))
        self.resnets = nn.ModuleList(rs)
        self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, ctx, groups) for _ in range(layers)]) if attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, temb, context):
        for i, r in enumerate(self.resnets):
            x = r(torch.cat([x, skips.pop()], dim=1), temb)
            if self.attentions is not None:
                x = self.attentions[i](x, context)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class CondMidBlock(nn.Module):
    def __init__(self, ch, temb_ch, groups, eps, heads, ctx):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb_ch, groups, eps) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(ch, heads, ctx, groups)])

    def forward(self, x, temb, context):
        return self.resnets[1](self.attentions[0](self.resnets[0](x, temb), context), temb)


class UNet2DConditionModel(nn.Module):
    def __init__(self, sample_size=64, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280),
                 layers_per_block=2, cross_attention_dim=768, attention_head_dim=8, norm_num_groups=32, norm_eps=1e-5,
                 down_block_types=None, up_block_types=None):
        super().__init__()
        n = len(block_out_channels)
        down_block_types = tuple(down_block_types or ("CrossAttnDownBlock2D",) * (n - 1) + ("DownBlock2D",))
        up_block_types = tuple(up_block_types or ("UpBlock2D",) + ("CrossAttnUpBlock2D",) * (n - 1))
        self.config = SimpleNamespace(sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
                                      block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                                      cross_attention_dim=cross_attention_dim, attention_head_dim=attention_head_dim,
                                      norm_num_groups=norm_num_groups, norm_eps=norm_eps,
                                      down_block_types=down_block_types, up_block_types=up_block_types)
        self.in_channels, self.sample_size = in_channels, sample_size
        boc = list(block_out_channels)
        temb_ch = boc[0] * 4
        g, eps, heads, ctx = norm_num_groups, norm_eps, attention_head_dim, cross_attention_dim   # SD 1.x: 8 = number of heads
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb_ch)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(CondDownBlock(in_ch, out_ch, temb_ch, layers_per_block, g, eps, t.startswith("CrossAttn"),
                                                  heads, ctx, i != n - 1))
        self.mid_block = CondMidBlock(boc[-1], temb_ch, g, eps, heads, ctx)
        self.up_blocks = nn.ModuleList()
        rev = boc[::-1]
        out_ch = rev[0]
        for i, t in enumerate(up_block_types):
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, n - 1)]
            self.up_blocks.append(CondUpBlock(in_ch, out_ch, prev, temb_ch, layers_per_block + 1, g, eps,
                                              t.startswith("CrossAttn"), heads, ctx, i != n - 1))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=eps)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states=None, **_):
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.long, device=sample.device)
        timestep = timestep.to(sample.device)
        if timestep.dim() == 0:
            timestep = timestep[None]
        timestep = timestep * torch.ones(sample.shape[0], dtype=timestep.dtype, device=sample.device)
        temb = timestep_embedding(timestep, self.config.block_out_channels[0], True, 0).to(sample.dtype)
        temb = self.time_embedding(temb)
        ctx = encoder_hidden_states
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, ctx)
            skips += outs
        x = self.mid_block(x, temb, ctx)
        for blk in self.up_blocks:
            x = blk(x, skips, temb, ctx)
        return UNet2DOutput(sample=self.conv_out(F.silu(self.conv_norm_out(x))))
