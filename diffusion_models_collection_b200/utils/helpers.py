"""Same public helpers as /root/reference/utils/helpers.py:12-133 (API kept so sample.py / train.py import them
from this package unchanged).  Trivial glue, not part of the accelerated path."""

from __future__ import annotations

import importlib.util
import json
import os
import random
import sys
from pathlib import Path

import numpy as np
import torch


def set_seed(seed=42):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def resolve_image_size(image_size):
    if isinstance(image_size, int):
        return (image_size, image_size)
    if isinstance(image_size, (list, tuple)) and len(image_size) == 2:
        h, w = image_size
        if not (isinstance(h, int) and isinstance(w, int)):
            raise ValueError("image_size values must be integers")
        return (h, w)
    raise ValueError("image_size must be int or a pair (H, W)")


def count_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def get_device(device_id=None):
    if device_id is not None:
        return torch.device(f"cuda:{device_id}")
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


def save_config(config, save_path):
    with Path(save_path).open("w", encoding="utf-8") as f:
        json.dump(config, f, indent=4)


def load_config(config_path):
    spec = importlib.util.spec_from_file_location("config", Path(config_path))
    module = importlib.util.module_from_spec(spec)
    sys.modules["config"] = module
    spec.loader.exec_module(module)
    return module.config


def normalize_to_neg_one_to_one(img):
    return img * 2 - 1


def unnormalize_to_zero_to_one(img):
    return (img + 1) * 0.5


def setup_distributed(rank, world_size, backend="nccl", port="12355"):
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(port))
    torch.distributed.init_process_group(backend, rank=rank, world_size=world_size)


def create_gif(images_list, save_path, fps=20):
    from PIL import Image

    frames = []
    for img in images_list:
        if isinstance(img, torch.Tensor):
            img = img.cpu().numpy()
        if img.ndim == 3 and img.shape[0] in (1, 3):
            img = np.transpose(img, (1, 2, 0))
        img = (img * 255).astype(np.uint8) if img.max() <= 1.0 else img.astype(np.uint8)
        if img.ndim == 3 and img.shape[2] == 1:
            img = img.squeeze(2)
        frames.append(Image.fromarray(img))
    frames[0].save(save_path, save_all=True, append_images=frames[1:], duration=1000 / fps, loop=0)
