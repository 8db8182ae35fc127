"""``diffusers.utils`` stub: only ``BaseOutput`` (src/SegDiffEditPipeline.py:7)."""
from oracle.ddim_scheduler import BaseOutput  # noqa: F401
