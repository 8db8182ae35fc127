#!/usr/bin/env python
"""Fused GroupNorm (conv_igemm XF kernels) against the stand-alone apply kernel: same arithmetic, so the noise prediction
must be BIT-IDENTICAL; prints the forward time of both.   python tools/check_gn_fuse.py [model] [batch]
Runs itself twice (B2E_GN_FUSE is read once per process)."""
import os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "diffusion-image-editing_b200"))


def child(model, B, out):
    import torch
    from models import create_diffusion_model
    w = create_diffusion_model(model, sample_clipping=False, max_batch=B, seed=0, with_encoder=False) if model != "ddpm" else \
        create_diffusion_model(model, sample_clipping=False, max_batch=B, seed=0)
    unet = w.unet
    cfg = unet.config
    x = torch.randn(B, cfg.in_channels, cfg.sample_size, cfg.sample_size, generator=torch.Generator().manual_seed(3)).cuda()
    kw = {}
    if model == "sd":
        kw["encoder_hidden_states"] = torch.randn(B, 77, 768, generator=torch.Generator().manual_seed(4)).cuda()
    for _ in range(3):
        eps = unet(x, 500, **kw)["sample"]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eps = unet(x, 500, **kw)["sample"]
    e1.record()
    torch.cuda.synchronize()
    print(f"  GN_FUSE={os.environ.get('B2E_GN_FUSE', '1')} {model} B={B}: {e0.elapsed_time(e1) / 10:.3f} ms / forward, "
          f"{unet.launches_per_forward} launches, finite={bool(torch.isfinite(eps).all())}", flush=True)
    outs = [eps.cpu()]
    if model in ("ldm", "sd"):
        dec = w.vqvae if model == "ldm" else w.vae
        z = torch.randn(min(B, dec.max_batch), dec.config.latent_channels, dec.config.sample_size, dec.config.sample_size,
                        generator=torch.Generator().manual_seed(5)).cuda().requires_grad_(True)
        img = dec.decode(z).sample
        g, = torch.autograd.grad(img.square().mean(), z)
        outs += [img.detach().cpu(), g.cpu()]
    torch.save(outs, out)


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[3] == "child":
        child(sys.argv[1], int(sys.argv[2]), sys.argv[4])
        sys.exit(0)
    import torch
    model = sys.argv[1] if len(sys.argv) > 1 else "ddpm"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    res = []
    for fuse in ("1", "0"):
        out = f"/tmp/gnfuse_{model}_{fuse}.pt"
        env = dict(os.environ, B2E_GN_FUSE=fuse)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), model, str(B), "child", out], env=env)
        if r.returncode:
            print(f"FAILED with B2E_GN_FUSE={fuse} rc={r.returncode}")
            sys.exit(1)
        res.append(torch.load(out))
    names = ["eps", "decoded image", "latent gradient"]
    ok = True
    for i, (a, b) in enumerate(zip(*res)):
        same = torch.equal(a, b)
        ok &= same
        print(f"  {names[i]}: bit-identical={same} max|diff|={(a - b).abs().max().item():.3e} max|ref|={b.abs().max().item():.3e}")
    print("GN_FUSE_OK" if ok else "GN_FUSE_MISMATCH")
