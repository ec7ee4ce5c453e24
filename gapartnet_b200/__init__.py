"""gapartnet_b200 - Blackwell (sm_100a) engine behind GAPartNet's spconv / epic_ops / pointnet2
operator surface.  See DESIGN.md; the C ABI is include/gapart_b200.h.

Importing the package needs no GPU; the CUDA library is loaded on first use and every op raises
if it (or CUDA) is missing - there is no CPU fallback.
"""
__version__ = "0.1.0"
