"""Per-op CUDA-event profile of the SD-1.x UNet forward:  python tools/profile_sd.py [batch] [out.json]"""
import sys, json, torch
sys.path.insert(0, "diffusion-image-editing_b200")
from b200edit.unet_cond import SD15_CONFIG, UNet2DConditionModel
from b200edit.unet import UNet2DModel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
OUT = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/sd_profile.json"
net = UNet2DConditionModel(**SD15_CONFIG, max_batch=B).init_random(0)
x = torch.randn(B, 4, 64, 64, device="cuda"); ctx = torch.randn(B, 77, 768, device="cuda")
net(x, 500, encoder_hidden_states=ctx)
prof = UNet2DModel.profile(net, x, 500)
by = {}
for p in prof:
    k = p["desc"].split(" ")[0] if p["kind"] != "conv_igemm" else ("attention gemm" if "attention" in p["desc"] else "conv/linear")
    if p["kind"] == "groupnorm": k = p["desc"] or "groupnorm"
    by.setdefault(k, [0, 0.0, 0.0]); by[k][0] += 1; by[k][1] += p["ms"]; by[k][2] += p["flops"]
for k, v in sorted(by.items(), key=lambda kv: -kv[1][1]): print(f"{k:28s} n={v[0]:4d} {v[1]:8.3f} ms  {v[2]/max(v[1],1e-9)/1e9:8.1f} TF/s")
print("total", sum(v[1] for v in by.values()))
json.dump(prof, open(OUT, "w"))
