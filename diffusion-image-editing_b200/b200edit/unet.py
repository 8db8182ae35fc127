"""Native UNet2DModel: same call contract as ``diffusers.UNet2DModel`` on the reference's path
(``unet(sample, timestep)["sample"]``, ``unet.config.in_channels/sample_size`` and the deprecated
direct attributes - src/diffusion_utils.py:72, src/utils.py:68-70, src/ddpm_inversion.py:39-41),
executed by libb200edit.so: bf16 tcgen05 implicit-GEMM convolutions, fp32 accumulation."""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import _C
from ._C import UNetConfig, check, lib

DDPM256_CONFIG = dict(
    sample_size=256, in_channels=3, out_channels=3,
    block_out_channels=(128, 128, 256, 256, 512, 512), layers_per_block=2,
    down_block_types=("DownBlock2D",) * 4 + ("AttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "AttnUpBlock2D") + ("UpBlock2D",) * 4,
    norm_num_groups=32, norm_eps=1e-6, attention_head_dim=None,
    flip_sin_to_cos=False, freq_shift=1, downsample_padding=0,
)

# CompVis/ldm-celebahq-256 `unet` (UNet2DModel on the 64x64x3 VQ latent, ~274 M parameters)
LDM_CELEBAHQ_CONFIG = dict(
    sample_size=64, in_channels=3, out_channels=3,
    block_out_channels=(224, 448, 672, 896), layers_per_block=2,
    down_block_types=("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"),
    up_block_types=("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
    norm_num_groups=32, norm_eps=1e-6, attention_head_dim=32,
    flip_sin_to_cos=True, freq_shift=0, downsample_padding=1,
)


class UNetOutput(dict):
    @property
    def sample(self):
        return self["sample"]


class UNet2DModel:
    def __init__(self, sample_size=256, in_channels=3, out_channels=3,
                 block_out_channels=(128, 128, 256, 256, 512, 512), layers_per_block=2,
                 down_block_types=None, up_block_types=None, norm_num_groups=32, norm_eps=1e-6,
                 attention_head_dim=None, flip_sin_to_cos=False, freq_shift=1, downsample_padding=0,
                 max_batch=8, device="cuda", precision=None):
        """precision: None / "fp16" (default: fp16 operands, fp32 accumulation) or "fp32" (fp32-accurate mode: split-f16
        operands hi + lo on the same tcgen05 kernels, three products per GEMM; meets the 1e-4 bar of the fp32
        reference at about a third of the fp16 throughput)."""
        _C.require_device()
        self.precision, self._precision_cfg = _C.resolve_precision(precision, "UNet2DModel")
        n = len(block_out_channels)
        down_block_types = tuple(down_block_types or ("DownBlock2D",) * n)
        up_block_types = tuple(up_block_types or ("UpBlock2D",) * n)
        self.config = SimpleNamespace(
            sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
            block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
            down_block_types=down_block_types, up_block_types=up_block_types,
            norm_num_groups=norm_num_groups, norm_eps=norm_eps, attention_head_dim=attention_head_dim,
            flip_sin_to_cos=flip_sin_to_cos, freq_shift=freq_shift, downsample_padding=downsample_padding)
        self.in_channels, self.sample_size = in_channels, sample_size   # deprecated aliases
        self.device = torch.device(device)
        self.dtype = torch.float32
        self.max_batch = int(max_batch)
        cfg = UNetConfig()
        cfg.sample_size, cfg.in_channels, cfg.out_channels, cfg.n_blocks = sample_size, in_channels, out_channels, n
        for i in range(n):
            cfg.block_out_channels[i] = block_out_channels[i]
            cfg.down_attn[i] = int(down_block_types[i].startswith("Attn"))
            cfg.up_attn[i] = int(up_block_types[i].startswith("Attn"))
        cfg.layers_per_block, cfg.norm_num_groups, cfg.norm_eps = layers_per_block, norm_num_groups, norm_eps
        cfg.attention_head_dim = int(attention_head_dim or 0)
        cfg.flip_sin_to_cos, cfg.freq_shift = int(flip_sin_to_cos), float(freq_shift)
        cfg.downsample_padding = int(downsample_padding)
        cfg.precision = self._precision_cfg
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.b2e_unet_create(C.byref(cfg), self.max_batch, C.byref(h)), "unet_create")
            self._h = h
            nbytes = lib.b2e_unet_workspace_bytes(h)
            self._ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            base = (self._ws.data_ptr() + 255) // 256 * 256
            check(lib.b2e_unet_bind_workspace(h, C.c_void_p(base), nbytes), "unet_bind_workspace")
        self._t_cache: Dict[tuple, torch.Tensor] = {}

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.b2e_unet_destroy(h)
            self._h = None

    # ------------------------------------------------------------------ parameters
    def param_info(self):
        out = []
        name, numel, fan = C.c_char_p(), C.c_int64(), C.c_int64()
        for i in range(lib.b2e_unet_num_params(self._h)):
            check(lib.b2e_unet_param_info(self._h, i, C.byref(name), C.byref(numel), C.byref(fan)))
            out.append((name.value.decode(), numel.value, fan.value))
        return out

    def load_state_dict(self, sd, strict=True):
        """sd: mapping diffusers-name -> fp32 tensor (any device)."""
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        names = {n for n, _, _ in self.param_info()}
        missing = sorted(names - set(sd))
        unexpected = sorted(set(sd) - names)
        if strict and (missing or unexpected):
            raise KeyError(f"load_state_dict: missing {missing[:5]}..., unexpected {unexpected[:5]}...")
        with torch.cuda.device(self.device):
            for k in names & set(sd):
                t = sd[k].detach().to(self.device, torch.float32).contiguous()
                check(lib.b2e_unet_set_param(self._h, k.encode(), C.c_void_p(t.data_ptr()), t.numel(), stream),
                      f"set_param({k})")
            torch.cuda.current_stream().synchronize()   # temporaries above must outlive the copies
        return missing, unexpected

    def init_random(self, seed=0):
        """PyTorch-default-style init (U(-1/sqrt(fan_in), 1/sqrt(fan_in)); norm scale 1, shift 0)."""
        self.load_state_dict(self.random_state_dict(seed))
        return self

    def random_state_dict(self, seed=0):
        """The state dict ``init_random(seed)`` loads (so that a second engine can be given the same weights)."""
        g = torch.Generator().manual_seed(seed)
        sd = {}
        for name, numel, fan in self.param_info():
            if fan == 0:
                sd[name] = torch.ones(numel) if name.endswith(".weight") else torch.zeros(numel)
            elif fan < 0:   # codebook: U(-1/n, 1/n) with n = -fan (diffusers VectorQuantizer init)
                sd[name] = (torch.rand(numel, generator=g) * 2 - 1) / float(-fan)
            else:
                bound = 1.0 / math.sqrt(fan)
                sd[name] = (torch.rand(numel, generator=g) * 2 - 1) * bound
        return sd

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    @property
    def flops_per_sample(self) -> float:
        return float(lib.b2e_unet_flops(self._h, 1))

    @property
    def launches_per_forward(self) -> int:
        return int(lib.b2e_unet_launches_per_forward(self._h))

    def profile(self, sample, timestep):
        """Per-op CUDA-event timings of one forward: list of dicts(kind, ms, flops, bytes)."""
        x = sample.to(torch.float32).contiguous()
        B = x.shape[0]
        t = self._timesteps(timestep, B)
        eps = torch.empty((B, self.config.out_channels, self.config.sample_size, self.config.sample_size),
                          dtype=torch.float32, device=x.device)
        cap = 1024
        n = C.c_int()
        ms, fl, by, kd = (C.c_float * cap)(), (C.c_double * cap)(), (C.c_double * cap)(), (C.c_int * cap)()
        check(lib.b2e_unet_profile(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(t.data_ptr()),
                                   C.c_void_p(eps.data_ptr()), B,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream), cap, C.byref(n), ms, fl, by, kd),
              "unet_profile")
        names = {0: "conv_igemm", 1: "groupnorm", 2: "attention", 3: "other"}
        return [dict(kind=names[kd[i]], ms=ms[i], flops=fl[i], bytes=by[i],
                     desc=lib.b2e_unet_op_desc(self._h, i).decode()) for i in range(n.value)]

    # ------------------------------------------------------------------ forward
    def _timesteps(self, timestep, B):
        if torch.is_tensor(timestep) and timestep.is_cuda:
            t = timestep.to(torch.int64)
            return (t.reshape(1).expand(B) if t.numel() == 1 else t.reshape(B)).contiguous()
        if torch.is_tensor(timestep) and timestep.numel() > 1:
            return timestep.to(self.device, torch.int64).contiguous()
        key = (int(timestep), B)
        if key not in self._t_cache:
            self._t_cache[key] = torch.full((B,), int(timestep), dtype=torch.int64, device=self.device)
        return self._t_cache[key]

    def __call__(self, sample, timestep, out: Optional[torch.Tensor] = None, **_):
        if not sample.is_cuda:
            raise _C.B2EError("UNet2DModel: sample must be a CUDA tensor (no CPU fallback)")
        x = sample.to(torch.float32).contiguous()
        B = x.shape[0]
        cfg = self.config
        if tuple(x.shape[1:]) != (cfg.in_channels, cfg.sample_size, cfg.sample_size):
            raise ValueError(f"UNet2DModel: expected (B,{cfg.in_channels},{cfg.sample_size},{cfg.sample_size}), "
                             f"got {tuple(x.shape)}")
        t = self._timesteps(timestep, B)
        eps = out if out is not None else torch.empty(
            (B, cfg.out_channels, cfg.sample_size, cfg.sample_size), dtype=torch.float32, device=x.device)
        if self._graph_on and B == self.max_batch and not torch.cuda.is_current_stream_capturing():
            return UNetOutput(sample=self._replay(x, t, eps))
        if self._graph_on and self._graphs:
            self._graphs.clear()      # an eager forward at another batch re-plans the workspace: captured launches are stale
        check(lib.b2e_unet_forward(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(t.data_ptr()),
                                   C.c_void_p(eps.data_ptr()), B,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)), "unet_forward")
        return UNetOutput(sample=eps)

    # ------------------------------------------------------------------ CUDA-graph replay (launch-bound regime)
    _graph_on = False

    def enable_cuda_graph(self, enable: bool = True):
        """Replay the forward's ~225 launches as ONE captured CUDA graph for calls at batch == max_batch (the launch-bound
        regime: at batch 1 the forward is 225 kernels of ~10 us).  The plan has fixed buffers, no allocation and no host
        synchronisation, so it captures as is (programmatic-dependent-launch edges included); inputs / timesteps / output go
        through static buffers.  Captured on first use; any eager call at another batch size re-plans the workspace and
        drops the capture."""
        self._graph_on = bool(enable)
        self._graphs = {}
        return self

    def _replay(self, x, t, eps):
        B = x.shape[0]
        g = self._graphs.get(B)
        if g is None:
            sx, st, se = torch.empty_like(x), torch.empty_like(t), torch.empty_like(eps)
            sx.copy_(x); st.copy_(t)
            stream = C.c_void_p
            # warm-up outside capture (plan rebuild for this batch, lazy function attributes), then capture
            check(lib.b2e_unet_forward(self._h, C.c_void_p(sx.data_ptr()), C.c_void_p(st.data_ptr()), C.c_void_p(se.data_ptr()), B,
                                       stream(torch.cuda.current_stream().cuda_stream)), "unet_forward")
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                check(lib.b2e_unet_forward(self._h, C.c_void_p(sx.data_ptr()), C.c_void_p(st.data_ptr()),
                                           C.c_void_p(se.data_ptr()), B, stream(torch.cuda.current_stream().cuda_stream)),
                      "unet_forward (capture)")
            g = self._graphs[B] = (graph, sx, st, se)
        graph, sx, st, se = g
        sx.copy_(x)
        st.copy_(t)
        graph.replay()
        eps.copy_(se)
        return eps
