from inpaintnet_b200.measure_vae import Encoder  # noqa: F401
