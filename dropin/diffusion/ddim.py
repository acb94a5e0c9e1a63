from diffusion_models_collection_b200.diffusion.ddim import DDIM  # noqa: F401
