from inpaintnet_b200.data import DatasetManager  # noqa: F401
