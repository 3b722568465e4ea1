from inpaintnet_b200.arnn import ConstraintModelGaussianReg, AnticipationRNNBaseline  # noqa: F401
from inpaintnet_b200.helpers import *  # noqa: F401,F403
