"""Drop-in for the reference's `pointnet2_cuda` extension and its `pointnet2_utils` autograd layer
(/root/reference/dataset/process_tools/utils/pointnet_lib/{src/pointnet2_api.cpp,pointnet2_utils.py})
and for `pointnet2_ops.pointnet2_utils.furthest_point_sample` (structure/utils.py:360)."""
from . import pointnet2_cuda, pointnet2_utils  # noqa: F401
