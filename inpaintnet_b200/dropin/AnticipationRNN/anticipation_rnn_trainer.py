from inpaintnet_b200.arnn import AnticipationRNNGaussianRegTrainer, AnticipationRNNBaselineTrainer  # noqa: F401
from inpaintnet_b200.helpers import *  # noqa: F401,F403
