from diffusion_models_collection_b200.models.unet import UNet  # noqa: F401,F403
