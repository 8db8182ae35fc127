#!/usr/bin/env python
"""One convolution for B2E_TRACE / B2E_DEBUG experiments: trace_conv.py N H Cin Cout [k]."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "diffusion-image-editing_b200"))
import torch
from b200edit import ops
N, H, cin, cout = (int(v) for v in sys.argv[1:5])
k = int(sys.argv[5]) if len(sys.argv) > 5 else 3
x = torch.randn(N, H, H, cin, device="cuda").to(ops.act_dtype())
w = torch.randn(cout, cin, k, k, device="cuda") / (cin * k * k) ** 0.5
b = torch.randn(cout, device="cuda")
for _ in range(4):
    out = ops.conv2d_nhwc_f16(x, w, b)
torch.cuda.synchronize()
