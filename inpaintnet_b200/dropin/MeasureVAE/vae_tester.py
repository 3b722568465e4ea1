from inpaintnet_b200.tester import VAETester  # noqa: F401
from inpaintnet_b200.helpers import *  # noqa: F401,F403
