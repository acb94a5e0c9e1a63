"""Denoiser models (drop-in for the reference's ``models`` package, models/__init__.py:6-10)."""
from .unet import UNet
from .dit import DiT
from .dim import DiM

__all__ = ["UNet", "DiT", "DiM"]
