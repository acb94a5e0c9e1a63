"""`from models import UNet, DiT, DiM` (reference: models/__init__.py:6-10) -> the B200-native denoisers."""
from diffusion_models_collection_b200.models import DiM, DiT, UNet  # noqa: F401

__all__ = ["UNet", "DiT", "DiM"]
