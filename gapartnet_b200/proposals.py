"""Static-capacity, sync-free proposal stage (gp_proposals_build, csrc/proposal.cu) behind a small tensor wrapper.

Reference: GAPartNet.proposal_clustering_and_revoxelize (/root/reference/gapartnet/network/model.py:228-346) and
segmented_voxelize (/root/reference/gapartnet/network/grouping_utils.py:47-104).  The reference's version builds its
outputs with boolean masks / unique_consecutive (data-dependent shapes, ~60 host syncs per step); here every output is a
preallocated buffer with a static capacity and the true sizes stay in `counts` on the device."""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import C, GapartError
from .ops import _p


class ProposalStage:
    """Buffers for N points in `batch` scenes: at most 2N proposal points and `max_proposals` proposals."""

    # counts[...] slots
    NV, NP, P, OVERFLOW = 0, 1, 2, 3

    def __init__(self, num_points: int, batch: int, max_proposals: int, device, *, radius: float, cap: int,
                 cap_shift: int, min_points: int, fullscale: float, scale_max: float):
        self.N, self.B, self.maxP = int(num_points), int(batch), int(max_proposals)
        self.radius, self.cap, self.cap_shift, self.min_points = float(radius), int(cap), int(cap_shift), int(min_points)
        self.fullscale, self.scale_max = float(fullscale), float(scale_max)
        nb = int(C.gp_proposals_ws_bytes(self.N, self.B, self.maxP))
        if nb <= 0:
            raise GapartError("gp_proposals_ws_bytes: bad sizes")
        i32 = dict(dtype=torch.int32, device=device)
        self.ws = torch.empty(nb + 256, dtype=torch.uint8, device=device)
        self._ws_off = (-self.ws.data_ptr()) % 256
        self._ws_bytes = nb
        self.counts = torch.zeros(8, **i32)
        self.v2o = torch.zeros(self.N, **i32)
        # index buffers start out as zeros: rows beyond the device counts always hold VALID (if stale) indices, so
        # static-shape gathers over the full capacity never read out of bounds
        self.sorted_indices = torch.zeros(2 * self.N + 1, **i32)
        self.prop_point = torch.zeros(2 * self.N + 1, **i32)
        self.proposal_indices = torch.zeros(2 * self.N + 1, **i32)
        self.proposal_offsets = torch.zeros(self.maxP + 1, dtype=torch.int64, device=device)
        self.sxyz = torch.zeros(2 * self.N, 3, dtype=torch.float32, device=device)
        self.range_min = torch.zeros(3, dtype=torch.float32, device=device)
        self.range_max = torch.full((3,), self.fullscale, dtype=torch.float32, device=device)

    def build(self, points: torch.Tensor, sem_preds: torch.Tensor, offsets: torch.Tensor,
              instance_labels: Optional[torch.Tensor], batch_offsets: torch.Tensor, rand: torch.Tensor):
        """points [N, >=3] fp32 (xyz first), sem_preds [N] int64, offsets [N,3] fp32, instance_labels [N] int32 or
        None, batch_offsets [B+1] int64, rand [2,3] fp32 (the two torch.rand(3) draws of grouping_utils.py:86-90)."""
        if points.shape[0] != self.N or sem_preds.dtype != torch.int64 or not offsets.is_contiguous():
            raise GapartError("ProposalStage.build: static capacity / dtype mismatch")
        if instance_labels is not None and instance_labels.dtype != torch.int32:
            raise GapartError("ProposalStage.build: instance_labels must be int32")
        C.gp_proposals_build(_p(points), points.stride(0), _p(sem_preds), _p(offsets), _p(instance_labels),
                             _p(batch_offsets), self.B, self.N, self.radius, self.cap, self.cap_shift, self.min_points,
                             self.fullscale, self.scale_max, _p(rand), self.maxP, self.ws.data_ptr() + self._ws_off,
                             self._ws_bytes, _p(self.counts), _p(self.v2o), _p(self.sorted_indices), _p(self.prop_point),
                             _p(self.proposal_indices), _p(self.proposal_offsets), _p(self.sxyz),
                             torch.cuda.current_stream().cuda_stream)

    def host_counts(self):
        """(Nv, Np, P) on the host - synchronises; diagnostics / tests only"""
        c = self.counts.tolist()
        if c[self.OVERFLOW]:
            raise GapartError(f"{c[self.OVERFLOW]} proposals exceed max_proposals={self.maxP}: the excess was dropped")
        return c[self.NV], c[self.NP], c[self.P]
