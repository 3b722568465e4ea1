from inpaintnet_b200.tester import LatentRNNTester  # noqa: F401
