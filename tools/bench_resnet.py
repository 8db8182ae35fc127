#!/usr/bin/env python
"""Forward + input-gradient time of the native ResNet-50 predictor vs torchvision (fp32 eager, cuDNN) on the same GPU:
python tools/bench_resnet.py [input_size] [batch]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "diffusion-image-editing_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402
import torchvision  # noqa: E402

from b200edit.resnet import resnet50_predictor  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
net = resnet50_predictor(80, S, max_batch=B)
ref = torchvision.models.resnet50()
ref.fc = torch.nn.Linear(2048, 80)
ref = ref.cuda().eval()
x = (torch.rand(B, 3, S, S, device="cuda") * 2 - 1)


def step(model):
    xx = x.clone().requires_grad_(True)
    out = model(xx)
    g, = torch.autograd.grad(out.view(-1, 40, 2)[0][31][0], xx)
    return g


def timeit(model, n=10):
    for _ in range(3):
        step(model)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step(model)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


tn, tr = timeit(net), timeit(ref)
fl = net.flops_per_sample * B
print(f"resnet50 {S}x{S} B={B}: native fwd+dgrad {tn:.3f} ms ({fl / tn / 1e9:.0f} TFLOP/s as executed, {fl / B / 1e9:.1f} GFLOP/img) | "
      f"torchvision fp32 eager fwd+autograd {tr:.3f} ms | x{tr / tn:.1f}")
