from inpaintnet_b200.model_base import Model  # noqa: F401
import os  # noqa: F401
import torch  # noqa: F401
