"""BASELINE.json configs 3 and 5: class-conditional CIFAR-10 UNet (10 classes + null label), 32x32."""
from diffusion_models_collection_b200.configs._base import UNET_PARAMS, make_config

config = make_config("unet", UNET_PARAMS, "cifar10-unet-cond", dataset="cifar10", conditional=True, num_classes=10,
                     cfg_scale=3.0)
