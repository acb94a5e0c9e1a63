"""Diffusion processes (drop-in for the reference's ``diffusion`` package, diffusion/__init__.py:6-9)."""
from .ddpm import DDPM
from .ddim import DDIM

__all__ = ["DDPM", "DDIM"]
