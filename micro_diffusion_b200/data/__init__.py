from .mds import LatentsDataset, MDSShard, build_streaming_latents_dataloader, write_mds  # noqa: F401
from .device_loader import DeviceBatchLoader  # noqa: F401
