"""Glue helpers with the reference's names (utils/helpers.py); nothing here imports swanlab."""
from .helpers import (count_parameters, create_gif, get_device, load_config, normalize_to_neg_one_to_one,
                      resolve_image_size, save_config, set_seed, setup_distributed, unnormalize_to_zero_to_one)
