from inpaintnet_b200.measure_vae import *  # noqa: F401,F403
from inpaintnet_b200.measure_vae import MeasureVAE, Encoder, HierarchicalDecoder, Decoder  # noqa: F401
from inpaintnet_b200.helpers import *  # noqa: F401,F403
