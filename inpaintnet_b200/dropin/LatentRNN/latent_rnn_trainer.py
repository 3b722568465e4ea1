from inpaintnet_b200.trainer import LatentRNNTrainer  # noqa: F401
