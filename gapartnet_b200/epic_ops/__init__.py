"""Drop-in for the `epic_ops` package as GAPartNet uses it
(/root/reference/gapartnet/network/grouping_utils.py:4-8, network/model.py, dataset/gapartnet.py:11)."""
from . import ball_query, ccl, iou, nms, reduce, voxelize  # noqa: F401
