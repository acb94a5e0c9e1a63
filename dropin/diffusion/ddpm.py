from diffusion_models_collection_b200.diffusion.ddpm import DDPM  # noqa: F401
