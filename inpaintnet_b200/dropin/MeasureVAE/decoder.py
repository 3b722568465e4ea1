from inpaintnet_b200.measure_vae import Decoder, HierarchicalDecoder  # noqa: F401
