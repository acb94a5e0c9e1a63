"""Builds config dicts in the reference's format (configs/cifar10_unet.py, configs/cifar10_dit.py: a module-level
``config`` dict read by utils.helpers.load_config and stored inside checkpoints).  Keys and defaults are the
reference's; the CIFAR-10 variants below are the ones BASELINE.json's configs name (32x32, 10 classes + null)."""


def make_config(model_type, model_params, experiment, *, dataset="custom", image_size=(32, 32), conditional=False,
                num_classes=102, epochs=200, batch_size=128, learning_rate=2e-4, data_root="./data", **overrides):
    cfg = {
        "project_name": "diffusion-models",
        "experiment_name": experiment,
        "model_type": model_type,
        "model_params": dict(model_params),
        "dataset": dataset,
        "data_root": data_root,
        "image_size": tuple(image_size),
        "conditional": conditional,
        "num_classes": num_classes,
        "use_subdirs": True,
        "label_file": None,
        "num_timesteps": 1000,
        "beta_start": 0.0001,
        "beta_end": 0.02,
        "beta_schedule": "linear",
        "loss_type": "l2",
        "cfg_scale": 1.3,
        "num_inference_steps": 50,
        "ddim_eta": 0.0,
        "epochs": epochs,
        "batch_size": batch_size,
        "num_workers": 4,
        "optimizer": "adamw",
        "learning_rate": learning_rate,
        "weight_decay": 1e-4,
        "gradient_accumulation_steps": 1,
        "use_ema": True,
        "ema_decay": 0.9999,
        "cfg_dropout_prob": 0.2,
        "use_scheduler": True,
        "scheduler_type": "cosine",
        "warmup_epochs": 10,
        "warmup_start_factor": 0.01,
        "save_dir": "./checkpoints",
        "save_interval": 10,
        "resume_path": None,
        "sample_dir": "./generated_images",
        "sample_interval": 20,
        "sample_start_epoch": 200,
        "num_samples": 16,
        "use_swanlab": False,
        "gpu_ids": [0],
        "port": "12355",
        "seed": 42,
    }
    cfg.update(overrides)
    return cfg


UNET_PARAMS = {
    "image_size": (32, 32), "in_channels": 3, "model_channels": 128, "out_channels": 3, "num_res_blocks": 2,
    "attention_resolutions": (16, 8), "dropout": 0.1, "channel_mult": (1, 2, 2, 2), "use_attention": True,
}

DIT_PARAMS = {
    "img_size": (64, 64), "patch_size": 2, "in_channels": 3, "hidden_size": 384, "depth": 12, "num_heads": 6,
    "mlp_ratio": 4.0, "dropout": 0.1,
}
