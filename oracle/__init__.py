"""CPU oracle for the GAPartNet hot path - TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the shipped path (gapartnet_b200) never does and fails loudly when the CUDA
library is missing.

What is restated and from where (paths relative to /root/reference):
  voxelize.py   epic_ops.voxelize contract as used by gapartnet/dataset/gapartnet.py:179-205 and
                gapartnet/network/grouping_utils.py:47-104
  rulebook.py   spconv indice pairs for SubMConv3d k3 / SparseConv3d k2 s2 / SparseInverseConv3d k2
                (call sites gapartnet/network/backbone.py:25-28,74-77,87-90)
  spconv_cpu.py spconv.pytorch module surface on CPU torch (gather-mm-index_add) + dense
                torch.nn.functional.conv3d second opinion
  cluster.py    epic_ops ball_query / ccl / reduce / iou / nms call contracts
                (gapartnet/network/grouping_utils.py:108-140,221-245; network/model.py:348-385)
  losses.py     the heads and loss functions of gapartnet/network/model.py:160-226,396-462, network/losses.py and
                grouping_utils.py:14-43 (PINNED against the reference's own functions: tests/golden/losses.npz)
  pointnet2.c   serial C restatement of dataset/process_tools/utils/pointnet_lib/src/*_gpu.cu
  _ref/         the reference's OWN pointnet2 CUDA kernels compiled from where they lie
                (Makefile), callable on the GPU box as a second oracle for row a17

PARITY STATUS: the arithmetic of spconv / epic_ops lives in third-party packages that are neither
vendored under /root/reference nor installable here (unpinned versions, README.md:57-60), and the
reference has no tests or golden vectors => for those rows this oracle is "parity unpinned": it
follows the reference's call sites, the published spconv conv arithmetic, and is cross-checked
against torch.nn.functional.conv3d.  Pinned exceptions: misc/pose_fitting.py (imported from the
reference to generate tests/golden/pose_*.npz) and the pointnet2 kernels (oracle/_ref).
"""
