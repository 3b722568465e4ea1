from inpaintnet_b200.trainer import VAETrainer  # noqa: F401
