from inpaintnet_b200.helpers import *  # noqa: F401,F403
from inpaintnet_b200.helpers import to_cuda_variable, to_cuda_variable_long, to_numpy, init_hidden_lstm  # noqa: F401
import torch  # noqa: F401  (the reference star-imports torch through utils.helpers)
