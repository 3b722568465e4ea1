from inpaintnet_b200.trainer import Trainer, EarlyStopping  # noqa: F401
