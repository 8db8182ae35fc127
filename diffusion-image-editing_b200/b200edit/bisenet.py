"""Native face parser: the reference's ``BiSeNet`` (src/Segmentation/model.py:234-262 on the ResNet-18 of
src/Segmentation/resnet.py) on libb200edit.so - the network behind ``SegmentationModel`` (src/models.py:80-118), whose
parsing map feeds the mask path (``prepare_for_edit``, src/SegDiffEditPipeline.py:79-97), and ``NetAttrFunc.loss``
(src/attr_functions.py:213-219), which differentiates through it.  ``net(x)[0]`` = logits (B, n_classes, S, S) like the
reference's first output; after ``enable_grad()`` the call is an autograd node backed by the native input gradient.

Eval-mode BatchNorm is folded into the convolutions when the reference's ``state_dict`` is loaded; convolutions run on
the tcgen05 implicit-GEMM kernel with ReLU in the epilogue, the channel-attention branches (global pooling -> 1x1
convolutions -> sigmoid) as small fp32 kernels, the final align_corners bilinear upsampling as one kernel."""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import torch

from . import _C
from ._C import ResNetConfig, check, lib
from .unet import UNet2DModel


class _BiSeNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, model):
        ctx.model, ctx.shape = model, tuple(image.shape)
        out = model._forward(image)
        ctx.gen = model._fwd_gen      # the activations of THIS forward live in the engine's single workspace
        return out

    @staticmethod
    def backward(ctx, d_logits):
        m = ctx.model
        if m._fwd_gen != ctx.gen:
            raise _C.B2EError("BiSeNet backward: the engine ran another forward since this graph was built (its activations "
                              "were overwritten); differentiate before the next call on the same network")
        g = d_logits.to(torch.float32).contiguous()
        dx = torch.empty(ctx.shape, dtype=torch.float32, device=g.device)
        check(lib.b2e_resnet_backward(m._h, C.c_void_p(g.data_ptr()), C.c_void_p(dx.data_ptr()), ctx.shape[0],
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "bisenet_backward")
        return dx, None


def _fold(w, sd, bn, eps=1e-5):
    g, b = sd[bn + ".weight"].double(), sd[bn + ".bias"].double()
    mu, var = sd[bn + ".running_mean"].double(), sd[bn + ".running_var"].double()
    s = g / torch.sqrt(var + eps)
    return (w.double() * s.view(-1, 1, 1, 1)).float(), (b - mu * s).float()


class BiSeNet(UNet2DModel):
    def __init__(self, n_classes=19, input_size=512, max_batch=1, device="cuda", precision=None):
        """precision: None / "fp16" - f16 operands throughout; "fp32" - fp32-accurate FORWARD (split f16 operands), so
        the ReLU / max-pool masks the backward pass routes gradients with are those of the fp32 network (see
        b200edit.resnet.ResNet)."""
        _C.require_device()
        self.precision, precision_cfg = _C.resolve_precision(precision, "BiSeNet")
        self.config = SimpleNamespace(n_classes=n_classes, input_size=input_size, in_channels=3, sample_size=input_size,
                                      out_channels=n_classes)
        self.device = torch.device(device)
        self.dtype = torch.float32
        self.max_batch = int(max_batch)
        cfg = ResNetConfig()
        cfg.input_size, cfg.in_channels, cfg.bottleneck, cfg.width, cfg.num_classes, cfg.head = input_size, 3, 0, 64, n_classes, 1
        for i in range(4):
            cfg.layers[i] = 2
        cfg.precision = precision_cfg
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.b2e_resnet_create(C.byref(cfg), self.max_batch, C.byref(h)), "bisenet_create")
            self._h = h
            nbytes = lib.b2e_unet_workspace_bytes(h)
            self._ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            base = (self._ws.data_ptr() + 255) // 256 * 256
            check(lib.b2e_unet_bind_workspace(h, C.c_void_p(base), nbytes), "unet_bind_workspace")
        self._t_cache = {}
        self.differentiable = False

    def enable_grad(self, enable: bool = True):
        """Gradient mode: ``net(x)[0]`` becomes differentiable w.r.t. the image (native dgrad of the whole parser;
        batches <= max_batch).  Costs workspace: the backward pass has its own buffers."""
        with torch.cuda.device(self.device):
            check(lib.b2e_unet_enable_grad(self._h, int(enable)), "unet_enable_grad")
            nbytes = lib.b2e_unet_workspace_bytes(self._h)
            self._ws = None
            self._ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            base = (self._ws.data_ptr() + 255) // 256 * 256
            check(lib.b2e_unet_bind_workspace(self._h, C.c_void_p(base), nbytes), "unet_bind_workspace")
        self.differentiable = bool(enable)
        return self

    _fwd_gen = 0   # bumped by every forward: a backward over a stale workspace is detected (see _BiSeNetFn)

    def _forward(self, x):
        cfg = self.config
        self._fwd_gen += 1
        o = torch.empty((x.shape[0], cfg.n_classes, cfg.input_size, cfg.input_size), dtype=torch.float32, device=x.device)
        check(lib.b2e_unet_forward(self._h, C.c_void_p(x.data_ptr()), None, C.c_void_p(o.data_ptr()), x.shape[0],
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)), "bisenet_forward")
        return o

    @staticmethod
    def fold_reference_state_dict(sd, eps=1e-5):
        """The reference's BiSeNet ``state_dict`` -> the engine's parameters (BatchNorm folded; the auxiliary heads
        conv_out16 / conv_out32 are training-time outputs and are dropped)."""
        out = {}
        for k, w in sd.items():
            if not k.endswith(".weight") or w.dim() != 4 or k.startswith(("conv_out16.", "conv_out32.")):
                continue
            name = k[:-len(".weight")]
            if name.startswith("cp.resnet."):
                if name.endswith("downsample.0"):
                    bn = name[:-1] + "1"
                else:
                    head, last = name.rsplit("conv", 1)
                    bn = f"{head}bn{last}"
                out[name + ".weight"], out[name + ".bias"] = _fold(w, sd, bn, eps)
            elif name.endswith(".conv") and (name + ".weight") in sd and (name[:-len(".conv")] + ".bn.weight") in sd:
                # ConvBNReLU: <block>.conv + <block>.bn -> <block>
                blk = name[:-len(".conv")]
                fw, fb = _fold(w, sd, blk + ".bn", eps)
                if blk == "cp.conv_avg":         # acts on the pooled vector: an fp32 matrix
                    fw = fw.reshape(fw.shape[0], -1)
                out[blk + ".weight"], out[blk + ".bias"] = fw, fb
            elif name.endswith(".conv_atten"):
                fw, fb = _fold(w, sd, name[:-len("conv_atten")] + "bn_atten", eps)
                out[name + ".weight"], out[name + ".bias"] = fw.reshape(fw.shape[0], -1), fb
            elif name in ("ffm.conv1", "ffm.conv2"):
                out[name + ".weight"] = w.float().reshape(w.shape[0], -1)
            elif name == "conv_out.conv_out":
                out[name + ".weight"], out[name + ".bias"] = w.float(), torch.zeros(w.shape[0])
        return out

    def load_reference_state_dict(self, sd, eps=1e-5):
        return self.load_state_dict(self.fold_reference_state_dict(sd, eps))

    def __call__(self, image):
        if not image.is_cuda:
            raise _C.B2EError("BiSeNet: image must be a CUDA tensor (no CPU fallback)")
        cfg = self.config
        if tuple(image.shape[1:]) != (3, cfg.input_size, cfg.input_size):
            raise ValueError(f"BiSeNet: expected (B,3,{cfg.input_size},{cfg.input_size}), got {tuple(image.shape)}")
        if image.requires_grad and torch.is_grad_enabled():
            if not self.differentiable:
                raise _C.B2EError("BiSeNet: call enable_grad() before differentiating through the native face parser")
            if image.shape[0] > self.max_batch:
                raise ValueError(f"BiSeNet with gradient: batch {image.shape[0]} > max_batch {self.max_batch}")
            xx = image if (image.dtype == torch.float32 and image.is_contiguous()) else image.to(torch.float32).contiguous()
            return (_BiSeNetFn.apply(xx, self), None, None)
        x = image.detach().to(torch.float32).contiguous()
        outs = []
        for b0 in range(0, x.shape[0], self.max_batch):
            self._fwd_gen += 1
            xb = x[b0:b0 + self.max_batch]
            o = torch.empty((xb.shape[0], cfg.n_classes, cfg.input_size, cfg.input_size), dtype=torch.float32, device=x.device)
            check(lib.b2e_unet_forward(self._h, C.c_void_p(xb.data_ptr()), None, C.c_void_p(o.data_ptr()), xb.shape[0],
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), "bisenet_forward")
            outs.append(o)
        return (outs[0] if len(outs) == 1 else torch.cat(outs), None, None)

    def eval(self):
        return self


class MultiResBiSeNet:
    """ONE face parser for every square input resolution - the reference's BiSeNet is fully convolutional, and its
    workflow shares one ``SegmentationModel`` between mask creation (512x512, src/models.py:113-118) and ``NetAttrFunc``
    (the decoded 256x256 image goes straight into ``segmentation_model.net``, src/attr_functions.py:213-215).  The
    engine plans its buffers per resolution, so this wrapper keeps one engine per input size (built on first use, same
    weights); differentiated calls get their own engine with the native input gradient enabled.

    ``precision``: None - forward-only engines (mask creation: an argmax) run f16 operands, the differentiated engines
    the fp32-accurate forward ("fp32"), so the guidance gradient has the fp32 network's ReLU / max-pool masks;
    "fp16" / "fp32" force one mode for both."""

    def __init__(self, n_classes=19, max_batch=1, device="cuda", seed=0, precision=None):
        self.n_classes, self.max_batch, self.device, self.seed = n_classes, int(max_batch), device, seed
        if precision not in (None, "fp16", "fp32"):
            raise ValueError(f"MultiResBiSeNet: precision must be None, 'fp16' or 'fp32' (got {precision!r})")
        self.precision = precision
        self._folded = None      # engine-format state dict (BatchNorm folded); None -> seeded random init
        self._engines = {}

    def load_reference_state_dict(self, sd, eps=1e-5):
        self._folded = BiSeNet.fold_reference_state_dict(sd, eps)
        for e in self._engines.values():
            e.load_state_dict(self._folded)
        return self

    def engine(self, size: int, grad: bool = False) -> BiSeNet:
        grad = bool(grad)
        e = self._engines.get((size, grad))
        if e is None:
            if size % 32 or size < 64:
                raise ValueError(f"BiSeNet: input resolution must be a multiple of 32 and >= 64 (got {size})")
            precision = self.precision or ("fp32" if grad else "fp16")
            e = BiSeNet(self.n_classes, size, max_batch=self.max_batch, device=self.device, precision=precision)
            if self._folded is not None:
                e.load_state_dict(self._folded)
            else:
                e.init_random(self.seed)
            if grad:
                e.enable_grad()
            self._engines[(size, grad)] = e
        return e

    def __call__(self, image):
        if image.dim() != 4 or image.shape[1] != 3 or image.shape[-1] != image.shape[-2]:
            raise ValueError(f"BiSeNet: expected (B,3,S,S), got {tuple(image.shape)}")
        return self.engine(int(image.shape[-1]), grad=image.requires_grad and torch.is_grad_enabled())(image)

    def eval(self):
        return self

    def to(self, *a, **k):
        return self
