from inpaintnet_b200.data import SyntheticFolkDataset as FolkDataset  # noqa: F401
from inpaintnet_b200.data import SyntheticFolkDataset as FolkDatasetNBars  # noqa: F401
from inpaintnet_b200.data import SyntheticFolkDataset as FolkMeasuresDataset  # noqa: F401
