"""In-step GPU data path: the per-sample CPU work of the reference's GAPartNetDataset.__getitem__
(/root/reference/gapartnet/dataset/gapartnet.py:66-82) on the whole batch on the device."""
from .gpu_prep import apply_augmentations, compact_instance_labels, draw_augmentation, generate_inst_info, prepare_batch  # noqa: F401
from . import prep  # noqa: F401  (offline frame preparation + packed shard format)
