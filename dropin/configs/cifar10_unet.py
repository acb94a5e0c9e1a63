from diffusion_models_collection_b200.configs.cifar10_unet import config  # noqa: F401
