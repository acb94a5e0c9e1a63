from diffusion_models_collection_b200.models.dit import DiT  # noqa: F401,F403
