"""CPU: the C-ABI library builds/loads and exports every symbol include/b200edit.h declares;
argument validation that needs no GPU."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(REPO, "include", "b200edit.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2e_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from b200edit import _C
    syms = declared_symbols()
    assert len(syms) >= 30
    raw = ctypes.CDLL(_C.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/b200edit.h but not exported"
    # and the python binding prototypes the same set
    assert sorted(_C.PROTOTYPES) == syms
    assert _C.lib.b2e_version() == 100


def test_struct_layouts_match_header():
    from b200edit import _C
    assert ctypes.sizeof(_C.StepCoeffs) == 9 * 4
    assert ctypes.sizeof(_C.GuidedStepParams) == 36 + 4 * 5 + 16 * 3 + 12
    assert ctypes.sizeof(_C.L2RegParams) == ctypes.sizeof(_C.GuidedStepParams) + 8
    assert ctypes.sizeof(_C.UNetConfig) == 4 * 4 + 32 * 3 + 4 * 10
    assert ctypes.sizeof(_C.ColorGradParams) == 16 * 3 + 4 * 3


def test_host_side_argument_validation():
    from b200edit import _C
    lib = _C.lib
    out = _C.StepCoeffs()
    assert lib.b2e_step_coeffs_compute(None, 0, 1.0, 0, 0, 0.0, 0, ctypes.byref(out)) == -1
    assert b"step_coeffs" in lib.b2e_last_error()
    p = _C.GuidedStepParams()
    assert lib.b2e_guided_step_f32(None, None, None, None, None, None, 1, 3, 16, ctypes.byref(p), None) == -1
    assert lib.b2e_mask_workspace_bytes(512, 512, 64, 64) > 512 * 512 * 4
    assert lib.b2e_mask_workspace_bytes(0, 512, 64, 64) == 0


def test_ops_refuse_cpu_tensors():
    import torch
    from b200edit import ops
    from b200edit._C import B2EError, StepCoeffs
    x = torch.zeros(1, 3, 4, 4)
    with pytest.raises(B2EError):
        ops.guided_step(x, x, StepCoeffs())
    with pytest.raises(B2EError):
        ops.mask_from_seg(torch.zeros(8, 8, dtype=torch.int64), [1], False, (4, 4))


def test_c_coefficients_within_one_ulp_of_torch():
    import numpy as np
    from b200edit import ops
    from oracle.ddim_scheduler import DDIMScheduler
    worst = 0.0
    for preset in ("ddpm", "ldm", "sd"):
        s = DDIMScheduler.from_preset(preset)
        s.set_timesteps(50)
        for t in s.timesteps.tolist():
            for eta in (0.0, 0.7, 1.0):
                for mode in ("ddim", "ddpm"):
                    a = ops.step_coeffs(s.alphas_cumprod, s.final_alpha_cumprod, t, t - 20, eta, mode)
                    b = ops.step_coeffs_c(s.alphas_cumprod, float(s.final_alpha_cumprod), t, t - 20, eta, mode)
                    for f, _ in a._fields_:
                        x, y = getattr(a, f), getattr(b, f)
                        if x == x and x != 0:
                            worst = max(worst, abs(x - y) / float(np.spacing(np.float32(abs(x)))))
    assert worst <= 1.0


def test_precision_names_follow_the_build():
    """The library is built for ONE 16-bit operand type (fp16 unless -DB2E_ACT_BF16); asking for the other raises instead of
    silently computing in a different precision."""
    from b200edit import _C
    fast = _C.fast_precision()
    assert fast == ("bf16" if _C.lib.b2e_act_dtype() == 1 else "fp16")
    assert _C.resolve_precision(None, "x") == (fast, 0)
    assert _C.resolve_precision(fast, "x") == (fast, 0)
    assert _C.resolve_precision("fp32", "x") == ("fp32", 1)
    other = "bf16" if fast == "fp16" else "fp16"
    with pytest.raises(ValueError):
        _C.resolve_precision(other, "x")
    with pytest.raises(ValueError):
        _C.resolve_precision("fp8", "x")


def test_new_guidance_entry_points_validate_arguments():
    from b200edit import _C
    lib = _C.lib
    p = _C.ColorGradParams()
    assert lib.b2e_color_loss_grad_f32(None, None, None, None, 1, 3, 16, ctypes.byref(p), None, 0, None) == -1
    assert b"color_loss_grad" in lib.b2e_last_error()
    assert lib.b2e_apply_latent_guidance_f32(None, None, None, 1, 48, 0, 1.0, 1.0, 1.0, None) == -1
    assert lib.b2e_color_loss_grad_workspace_bytes() >= 8 * 148
