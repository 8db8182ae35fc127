"""Model wrappers that define the image <-> latent maps (drop-in for src/diffusion_classes.py)."""
import torch

from b200edit._C import B2EError

from base_diffusion import Diffusion
from diffusion_utils import prep_text


class DDPM(Diffusion):
    """Pixel-space diffusion: encode/decode are the identity, which is what lets the guidance
    gradient be evaluated analytically inside the fused step kernel."""

    decode_is_identity = True

    def encode(self, sample: torch.Tensor, **kwargs) -> torch.Tensor:
        return sample

    def decode(self, latent: torch.Tensor, **kwargs) -> torch.Tensor:
        return latent


class LDM(Diffusion):
    """Latent diffusion with a VQ-VAE (``model.vqvae`` must provide encode().latents / decode().sample)."""

    def __init__(self, model):
        super().__init__(model)
        self.vqvae = model.vqvae
        # optional differentiable twin of a native (forward-only) vqvae, used when the guidance graph runs
        # through the decoder (AttrFunc.apply -> decode(no_grad=False), src/attr_functions.py:153)
        self.guidance_vqvae = getattr(model, "guidance_vqvae", None)

    def native_decoder(self):
        """(decoder engine, chain factor d(decoder input)/d(x0 prediction)) when the guidance graph can run on the native
        decoder without autograd (analytic colour guidance, AttrFunc.apply); None otherwise."""
        for vq in (self.vqvae, self.guidance_vqvae):     # the twin: the gradient engine of an fp32-accurate pipeline
            if vq is not None and hasattr(vq, "decode_keep") and not getattr(vq, "forward_only", True):
                return vq, 1.0
        return None

    def encode(self, sample: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            return self.vqvae.encode(sample.to(dtype=torch.float32)).latents

    def decode(self, latent: torch.Tensor, no_grad=True) -> torch.Tensor:
        latent = latent.to(dtype=torch.float32)
        if no_grad:
            with torch.no_grad():
                return self.vqvae.decode(latent).sample
        # guidance graph: a native VQModel in gradient mode is itself an autograd node (native dgrad)
        vq = self.vqvae
        if getattr(vq, "forward_only", False):
            if self.guidance_vqvae is None:
                raise B2EError("LDM.decode(no_grad=False): the native VQ decoder is forward-only; pass a "
                               "differentiable module as guidance_vqvae= for guidance through the decoder")
            vq = self.guidance_vqvae
        return vq.decode(latent).sample


class SD(Diffusion):
    """Stable Diffusion: KL-VAE latents scaled by 0.18215, CLIP text conditioning."""

    SCALE = 0.18215

    def __init__(self, model):
        super().__init__(model)
        self.vae = model.vae
        # optional gradient twin of a forward-only native vae (fp32-accurate pipelines: the split-operand decoder has no
        # backward pass; guidance runs through an f16-operand engine with the same weights, gradient within 3e-3 of fp32)
        self.guidance_vae = getattr(model, "guidance_vae", None)
        self.tokenizer = model.tokenizer
        self.text_encoder = model.text_encoder

    def native_decoder(self):
        """See LDM.native_decoder; SD.decode feeds the decoder latent / 0.18215."""
        for vae in (self.vae, self.guidance_vae):
            if vae is not None and hasattr(vae, "decode_keep") and not getattr(vae, "forward_only", True):
                return vae, 1.0 / self.SCALE
        return None

    def encode(self, sample: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            latent = self.vae.encode(sample).latent_dist.mode().detach()
        return self.SCALE * latent

    def decode(self, latent: torch.Tensor, no_grad=True) -> torch.Tensor:
        latent = 1 / self.SCALE * latent
        if no_grad:
            with torch.no_grad():
                return self.vae.decode(latent).sample
        vae = self.vae
        if getattr(vae, "forward_only", False) and self.guidance_vae is not None:
            vae = self.guidance_vae
        return vae.decode(latent).sample

    def additional_prep(self, model, prompt):
        return prep_text(model, prompt)
