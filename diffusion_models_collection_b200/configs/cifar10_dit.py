"""DiT config, same dict as the reference ships in configs/cifar10_dit.py (64x64 tokens-1024 variant)."""
from diffusion_models_collection_b200.configs._base import DIT_PARAMS, make_config

config = make_config("dit", DIT_PARAMS, "cifar10-dit-ddpm", image_size=(64, 64), epochs=2000, batch_size=16,
                     learning_rate=1e-4)
