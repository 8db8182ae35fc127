"""GPU numerics of the native CLIP text encoder (``text_encoder(ids)[0]`` of prep_text, src/diffusion_utils.py:34-52)
against transformers' CLIPTextModel - the reference's own dependency - in fp32 (torch eager on the GPU as the checker)
with the same random-init weights.  Tolerance (IEEE f16 operands / residual stream, fp32 accumulation): relative RMS of
the last hidden state <= 4e-3 (measured 0.8e-3 .. 1.3e-3) and no worse than 1.5x transformers itself run in 16 bit."""
import pytest
import torch

pytestmark = pytest.mark.gpu
transformers = pytest.importorskip("transformers")


def run_pair(cfg, B, L, seed):
    from b200edit.clip import CLIPTextModel
    from transformers import CLIPTextConfig, CLIPTextModel as RefCLIP
    torch.manual_seed(seed)
    ref = RefCLIP(CLIPTextConfig(hidden_act="quick_gelu", **cfg)).eval()
    native = CLIPTextModel(**cfg, max_batch=B)
    native.load_state_dict(ref.state_dict())
    ids = torch.randint(0, cfg["vocab_size"], (B, L), generator=torch.Generator().manual_seed(seed + 1))
    got = native(ids.cuda())[0]
    torch.cuda.synchronize()
    with torch.no_grad():
        torch.backends.cuda.matmul.allow_tf32 = False
        rc = ref.cuda()
        want = rc(ids.cuda())[0]
        want16 = rc.half()(ids.cuda())[0].float()
    return got, want, want16


def check(got, ref, ref16, tag):
    rel = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    rel16 = ((ref16 - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    err = (got - ref).abs().max().item()
    print(f"{tag}: last_hidden_state rel-rms native {rel:.3e} max-abs {err:.3e} | transformers-bf16 {rel16:.3e}")
    assert got.shape == ref.shape and torch.isfinite(got).all()
    assert rel <= 4e-3 and rel <= 1.5 * rel16 + 2e-4      # measured 0.8e-3 .. 1.3e-3 (bf16 build of round 1: 1.0e-2)


SMALL = dict(vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
             max_position_embeddings=77)


@pytest.mark.parametrize("B,L", [(1, 77), (3, 20)])
def test_small_clip_matches_transformers(B, L):
    check(*run_pair(SMALL, B, L, seed=B), f"small clip B={B} L={L}")


def test_sd15_clip_text_encoder_matches_transformers():
    """The full CLIP ViT-L/14 text tower of Stable Diffusion 1.x (123 M parameters), the [uncond, prompt] pair of prep_text."""
    from b200edit.clip import SD15_CLIP_CONFIG
    check(*run_pair(SD15_CLIP_CONFIG, 2, 77, seed=5), "sd-1.x clip text encoder")


def test_causality_and_prep_text_contract():
    """Token t's state does not depend on later tokens; prep_text concatenates the "" and prompt encodings -> (2, 77, D)."""
    from b200edit.clip import CLIPTextModel
    from diffusion_utils import prep_text
    from types import SimpleNamespace
    enc = CLIPTextModel(**SMALL, max_batch=2).init_random(3)
    ids = torch.randint(0, 1000, (1, 77), generator=torch.Generator().manual_seed(4)).cuda()
    ids2 = ids.clone()
    ids2[0, 40:] = 7
    a, b = enc(ids)[0], enc(ids2)[0]
    assert torch.equal(a[0, :40], b[0, :40]) and not torch.equal(a[0, 40:], b[0, 40:])

    class Tok:
        model_max_length = 77

        def __call__(self, prompts, **kw):
            out = torch.zeros(len(prompts), 77, dtype=torch.long)
            for i, p in enumerate(prompts):
                for j, ch in enumerate(p.encode()[:77]):
                    out[i, j] = 1 + ch % 97
            return SimpleNamespace(input_ids=out)

    model = SimpleNamespace(tokenizer=Tok(), text_encoder=enc, device=torch.device("cuda"))
    emb = prep_text(model, "a photo of a face")
    assert emb.shape == (2, 77, 128) and torch.isfinite(emb).all()
