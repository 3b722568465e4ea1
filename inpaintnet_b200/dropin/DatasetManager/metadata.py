from inpaintnet_b200.data import TickMetadata, BeatMarkerMetadata  # noqa: F401
