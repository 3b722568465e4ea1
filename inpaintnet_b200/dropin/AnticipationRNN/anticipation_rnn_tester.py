from inpaintnet_b200.tester import AnticipationRNNTester  # noqa: F401
from inpaintnet_b200.helpers import *  # noqa: F401,F403
