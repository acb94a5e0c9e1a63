from diffusion_models_collection_b200.models.dim import DiM  # noqa: F401,F403
