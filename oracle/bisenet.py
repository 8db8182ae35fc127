"""Oracle restatement of the reference's face parser ``BiSeNet`` (src/Segmentation/model.py:234-262 on the ResNet-18 of
src/Segmentation/resnet.py:58-80): the network behind ``SegmentationModel`` (src/models.py:80-118) and
``NetAttrFunc.loss`` (src/attr_functions.py:213-219).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PINNED: same module tree / state_dict names as the reference, checked
against the unmodified reference module (imported from /root/reference with its hub download stubbed) on seeded
weights - tests/golden/make_golden.py writes tests/golden/bisenet.npz, tests/test_oracle_golden.py checks it.
Only ``out[0]`` is restated (the auxiliary heads conv_out16 / conv_out32 are training-time outputs the path never reads)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class ConvNormAct(nn.Module):
    def __init__(self, cin, cout, ks=3, stride=1, padding=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, ks, stride, padding, bias=False)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class ResBlock2(nn.Module):
    def __init__(self, cin, cout, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if cin != cout or stride != 1:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        r = self.bn2(self.conv2(F.relu(self.bn1(self.conv1(x)))))
        return F.relu((x if self.downsample is None else self.downsample(x)) + r)


class Backbone18(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = nn.Sequential(ResBlock2(64, 64), ResBlock2(64, 64))
        self.layer2 = nn.Sequential(ResBlock2(64, 128, 2), ResBlock2(128, 128))
        self.layer3 = nn.Sequential(ResBlock2(128, 256, 2), ResBlock2(256, 256))
        self.layer4 = nn.Sequential(ResBlock2(256, 512, 2), ResBlock2(512, 512))

    def forward(self, x):
        x = self.layer1(self.maxpool(F.relu(self.bn1(self.conv1(x)))))
        f8 = self.layer2(x)
        f16 = self.layer3(f8)
        return f8, f16, self.layer4(f16)


class ChannelAttentionRefine(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = ConvNormAct(cin, cout)
        self.conv_atten = nn.Conv2d(cout, cout, 1, bias=False)
        self.bn_atten = nn.BatchNorm2d(cout)

    def forward(self, x):
        feat = self.conv(x)
        atten = torch.sigmoid(self.bn_atten(self.conv_atten(F.avg_pool2d(feat, feat.shape[2:]))))
        return feat * atten


class ContextBranch(nn.Module):
    def __init__(self):
        super().__init__()
        self.resnet = Backbone18()
        self.arm16 = ChannelAttentionRefine(256, 128)
        self.arm32 = ChannelAttentionRefine(512, 128)
        self.conv_head32 = ConvNormAct(128, 128)
        self.conv_head16 = ConvNormAct(128, 128)
        self.conv_avg = ConvNormAct(512, 128, ks=1, stride=1, padding=0)

    def forward(self, x):
        f8, f16, f32 = self.resnet(x)
        avg = self.conv_avg(F.avg_pool2d(f32, f32.shape[2:]))
        s32 = self.arm32(f32) + F.interpolate(avg, f32.shape[2:], mode="nearest")
        u32 = self.conv_head32(F.interpolate(s32, f16.shape[2:], mode="nearest"))
        s16 = self.arm16(f16) + u32
        u16 = self.conv_head16(F.interpolate(s16, f8.shape[2:], mode="nearest"))
        return f8, u16, u32


class FusionBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.convblk = ConvNormAct(cin, cout, ks=1, stride=1, padding=0)
        self.conv1 = nn.Conv2d(cout, cout // 4, 1, bias=False)
        self.conv2 = nn.Conv2d(cout // 4, cout, 1, bias=False)

    def forward(self, fsp, fcp):
        feat = self.convblk(torch.cat([fsp, fcp], dim=1))
        atten = torch.sigmoid(self.conv2(F.relu(self.conv1(F.avg_pool2d(feat, feat.shape[2:])))))
        return feat * atten + feat


class ParserHead(nn.Module):
    def __init__(self, cin, mid, n_classes):
        super().__init__()
        self.conv = ConvNormAct(cin, mid)
        self.conv_out = nn.Conv2d(mid, n_classes, 1, bias=False)

    def forward(self, x):
        return self.conv_out(self.conv(x))


class BiSeNet(nn.Module):
    def __init__(self, n_classes=19):
        super().__init__()
        self.cp = ContextBranch()
        self.ffm = FusionBlock(256, 256)
        self.conv_out = ParserHead(256, 256, n_classes)
        self.conv_out16 = ParserHead(128, 64, n_classes)
        self.conv_out32 = ParserHead(128, 64, n_classes)

    def forward(self, x):
        f8, cp8, cp16 = self.cp(x)
        out = self.conv_out(self.ffm(f8, cp8))
        out = F.interpolate(out, x.shape[2:], mode="bilinear", align_corners=True)
        return out, None, None


def randomize_batchnorm(net: nn.Module, seed: int):
    """Non-trivial eval-mode BatchNorm statistics / affine parameters (a trained network's are not (0, 1, 1, 0))."""
    g = torch.Generator().manual_seed(seed)
    for mod in net.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
            mod.weight.data.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
            mod.bias.data.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
    return net


def seeded_weights(net: nn.Module, seed: int):
    """Fill every parameter / buffer from one seeded stream in sorted state_dict order, independent of how the module
    tree initialises itself (the reference's constructors draw kaiming-normal weights in their own order), so that the
    reference module and this restatement get identical weights from the same seed."""
    g = torch.Generator().manual_seed(seed)
    sd = net.state_dict()
    for k in sorted(sd):
        v = sd[k]
        if k.endswith("num_batches_tracked"):
            continue
        if v.dim() == 4:
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            v.copy_(torch.randn(v.shape, generator=g) * (2.0 / fan_in) ** 0.5)
        elif k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.75)
        elif k.endswith("running_mean") or k.endswith(".bias"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        else:   # BatchNorm scale
            v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.75)
    return net
