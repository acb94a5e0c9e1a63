from diffusion_models_collection_b200.configs.cifar10_dit import config  # noqa: F401
