from diffusion_models_collection_b200.utils.helpers import *  # noqa: F401,F403
