#!/usr/bin/env python
"""BASELINE config 4 timing (per GPU): Stable Diffusion 1.x layout, CFG (doubled latent batch) + classifier guidance whose
gradient flows through the NATIVE ResNet-50 attribute predictor (or, with a third argument "torch", torchvision's module)
and the NATIVE KL decoder (forward + gradient).  One guided step = conditional UNet forward on 2B latents + CFG combine + fused DDIM step + decoder forward
(B x 3 x 512 x 512) + classifier forward / backward + decoder gradient + update.   python tools/bench_sd.py BATCH [steps] [native|torch]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "diffusion-image-editing_b200"))
import torch
from attr_functions import ClassifierAttrFunc
from models import create_diffusion_model
from SegDiffEditPipeline import SegDiffEditPipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3


class Tok:
    model_max_length = 77
    def __call__(self, prompts, **kw):
        from types import SimpleNamespace
        return SimpleNamespace(input_ids=torch.zeros(len(prompts), 77, dtype=torch.long))


class Enc(torch.nn.Module):   # stands in for CLIP (no weights offline): (1, 77) ids -> (1, 77, 768)
    def __init__(self):
        super().__init__()
        self.e = torch.nn.Embedding(8, 768)
    def forward(self, ids):
        return (self.e(ids),)


PRED = sys.argv[3] if len(sys.argv) > 3 else "native"
if PRED == "native":   # the attribute predictor on the engine (forward + input gradient on the tcgen05 kernels)
    from models import get_pretrained_anyGAN
    predictor = get_pretrained_anyGAN(input_size=512, max_batch=B)
else:                  # torchvision's module differentiated by torch autograd (cuDNN), the earlier configuration
    import torchvision
    predictor = torchvision.models.resnet50(num_classes=80).cuda().eval()
w = create_diffusion_model("sd", sample_clipping=False, max_batch=B, seed=0, tokenizer=Tok(), text_encoder=Enc().cuda())
w.scheduler.set_timesteps(K)
pipe = SegDiffEditPipeline(w, None)
xt = torch.randn(B, 4, 64, 64, generator=torch.Generator().manual_seed(4)).cuda()
f = ClassifierAttrFunc(predictor, idx_for_class=31, idx_of_interest=0, loss_scale=50.0, t1=0, t2=K)


def run():
    return pipe.edit_image(xt=xt, attr_func=f, prompt="a photo of a face", cfg_scale=7.5, prog_bar=False, output_type="tensor")


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
fl = (2 * w.unet.flops_per_sample + w.vae.flops_per_sample) * B
print(f"config 4, B={B}/GPU: {ms:.1f} ms per guided step, {B / ms * 1e3:.1f} guided img-steps/s, ~{fl / ms / 1e9:.0f} TFLOP/s native "
      f"(UNet 2 x {w.unet.flops_per_sample / 1e12:.3f} + KL decoder fwd+bwd {w.vae.flops_per_sample / 1e12:.3f} TFLOP/img; "
      f"classifier: {PRED}; one final decode per {K} steps)")
print("max memory GB", torch.cuda.max_memory_allocated() / 1e9)
