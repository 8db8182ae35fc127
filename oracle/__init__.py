"""CPU oracle for the guided denoising loop -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This package restates, in plain PyTorch / numpy on the CPU, the arithmetic of
the hot path of JohanLundberg12/diffusion-image-editing:

* ``oracle.ddim_scheduler``  - the slice of ``diffusers.DDIMScheduler`` the
  reference calls (third-party, NOT in /root/reference, version unpinned).
* ``oracle.unet2d``          - ``diffusers.UNet2DModel`` (DDPM-256 layout).
* ``oracle.step_math``       - reference-owned element-wise step math
  (src/diffusion_utils.py, src/ddpm_inversion.py, src/ddim_inversion.py,
  src/attr_functions.py, src/utils.py).
* ``oracle.mask``            - src/mask_creator.py + src/Morphology.py in numpy.
* ``oracle.loops``           - the loop compositions (edit_image / invert / sample).
* ``oracle.upsample_phases`` - the sub-pixel phase decomposition of ``Upsample2D`` (nearest x2 + conv3x3) the engine
  computes, restated and pinned against the plain composition (tests/test_upsample_phases_cpu.py).
* ``oracle/shims``           - stub packages named ``diffusers`` and ``lpips`` so
  that the UNMODIFIED reference ``src/*.py`` can be imported in the dev container
  (``tests/golden/make_golden.py``); they re-export the restatements above.

Pinning status
--------------
* Reference-owned arithmetic (step_math, mask, loops): PINNED against the
  reference itself.  ``tests/golden/make_golden.py`` imports the unmodified
  reference from /root/reference/src (on top of ``oracle/shims``) in the dev
  container, runs it on seeded inputs and on the reference's own fixture
  images, and commits the input/output vectors under ``tests/golden/``.
  ``tests/test_oracle_golden.py`` checks every oracle function against them.
* ``diffusers`` arithmetic (DDIMScheduler.step/set_timesteps/add_noise,
  UNet2DModel): PARITY UNPINNED.  diffusers is a third-party dependency that is
  neither vendored in the reference nor installed/installable here (no
  network), and the reference pins no version and ships no golden vectors.
  The restatement follows the published algorithm (DDIM eq. 12/16, the public
  UNet2DModel layout of google/ddpm-celebahq-256) and becomes the normative
  spec for this repo.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product path
(``diffusion-image-editing_b200/``) never does.
"""
