"""UNet config, same dict as the reference ships in configs/cifar10_unet.py (unconditional, 32x32)."""
from diffusion_models_collection_b200.configs._base import UNET_PARAMS, make_config

config = make_config("unet", UNET_PARAMS, "cifar10-unet-ddpm")
