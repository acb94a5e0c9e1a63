"""`from diffusion import DDPM, DDIM` (reference: diffusion/__init__.py:6-9) -> the fused-kernel samplers."""
from diffusion_models_collection_b200.diffusion import DDIM, DDPM  # noqa: F401

__all__ = ["DDPM", "DDIM"]
