#!/usr/bin/env python
"""Single big convolution (8 x 256 x 256, Cin -> 128, 3x3) for B2E_TRACE / B2E_DEBUG experiments."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "diffusion-image-editing_b200"))
import torch
from b200edit import ops
cin = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = 8
x = torch.randn(N, 256, 256, cin, device="cuda").bfloat16()
w = torch.randn(128, cin, 3, 3, device="cuda") / (cin * 9) ** 0.5
b = torch.randn(128, device="cuda")
for _ in range(3):
    out = ops.conv2d_nhwc_bf16(x, w, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
# the test hook repacks the weights on every call; time the whole call but report separately
e0.record()
for _ in range(5):
    out = ops.conv2d_nhwc_bf16(x, w, b)
e1.record()
torch.cuda.synchronize()
print("ms per call (incl. weight repack)", e0.elapsed_time(e1) / 5)
