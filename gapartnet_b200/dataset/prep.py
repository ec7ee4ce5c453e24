"""Offline preparation of rendered frames on the GPU + a packed, memory-mappable shard format.

The reference converts one rendered frame at a time on the CPU
(/root/reference/dataset/process_tools/convert_rendered_into_input.py): a Python double loop over the pixels
(`get_point_cloud` :40-66), farthest point sampling down to 20 000 points (:112, pointnet2's kernel through
utils/sample_utils.py:26-29), ball normalisation (:69-87), label conversion with a relabel loop (:125-142), and one
pickled `.pth` tuple per frame (:156) that `dataset/gapartnet.py:208-229` un-pickles again in every DataLoader worker.

Here a frame is back-projected, sampled (csrc/pointnet2.cu: gp_pn2_furthest_point_sampling, bit-exact with the
reference's kernel and 7.7x faster at 80 000 -> 20 000, profiles/pointnet2_vs_reference_r2.jsonl), normalised and
relabelled on the device, and frames are stored in ONE file per split with fixed-size sections that `numpy.memmap`
serves without un-pickling:

    offset 0    : magic "GAPSHRD1", u32 version, u32 num_frames, u32 num_points, u32 name_bytes, u32 capacity, pad to 64
    sections    : xyz f32[S,N,3] | rgb f32[S,N,3] | sem i32[S,N] | ins i32[S,N] | npcs f32[S,N,3] | idx i32[S,N,2] |
                  scale f64[S,4] (max_radius, center xyz: the reference's meta/*.txt) | names u8[S,name_bytes]
                  each section 64-byte aligned, frames contiguous inside a section

`ShardReader.pth_tuple(i)` is exactly the tuple the reference's `torch.load(file_path)` returns, `load_data(i)` the
fields its `load_data` builds from it (dataset/gapartnet.py:214-229).
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .._lib import C, GapartError
from ..ops import _stream

MAX_INSTANCE_NUM = 1000          # convert_rendered_into_input.py:33
MAGIC = b"GAPSHRD1"
_SECTIONS = (("xyz", np.float32, 3), ("rgb", np.float32, 3), ("sem", np.int32, 0), ("ins", np.int32, 0),
             ("npcs", np.float32, 3), ("idx", np.int32, 2))


# ---------------------------------------------------------------------------------------------- frame -> arrays (GPU)
def back_project(rgb_image: torch.Tensor, depth_map: torch.Tensor, sem_seg_map: torch.Tensor, ins_seg_map: torch.Tensor,
                 npcs_map: torch.Tensor, K: torch.Tensor):
    """get_point_cloud (:40-66) without the pixel loop: pixels in row-major order, those labelled -2 in either map are
    skipped.  Arithmetic in float64 like the reference's Python floats.  -> (pcs f64[M,3], rgb f64[M,3], sem, ins,
    npcs, idx i64[M,2] (y, x))"""
    H, W = depth_map.shape
    keep = ~((sem_seg_map == -2) | (ins_seg_map == -2))
    yx = keep.nonzero()                                   # row-major, like the loop
    y_, x_ = yx[:, 0], yx[:, 1]
    Kd = K.double()
    z = depth_map[y_, x_].double()
    x_new = (x_.double() - Kd[0, 2]) * z / Kd[0, 0]
    y_new = (y_.double() - Kd[1, 2]) * z / Kd[1, 1]
    pcs = torch.stack([x_new, y_new, z], 1)
    # a true division: torch turns `tensor / python_float` into a multiplication by the reciprocal, 1 ulp off numpy
    rgb = rgb_image[y_, x_].double() / torch.full((), 255.0, dtype=torch.float64, device=depth_map.device)
    return (pcs, rgb, sem_seg_map[y_, x_], ins_seg_map[y_, x_], npcs_map[y_, x_], yx)


def fps(pcs: torch.Tensor, num_points: int) -> torch.Tensor:
    """FPS (utils/sample_utils.py:46-66): float32 copy of the points, first sample = point 0 -> int64 [num_points]"""
    if not pcs.is_cuda:
        raise GapartError("dataset.prep.fps needs CUDA tensors (there is no CPU fallback)")
    n = pcs.shape[0]
    if n == num_points:
        return torch.arange(n, device=pcs.device)
    xyz = pcs.float().contiguous()
    temp = torch.full((n,), 1e10, dtype=torch.float32, device=pcs.device)
    idx = torch.empty(num_points, dtype=torch.int32, device=pcs.device)
    C.gp_pn2_furthest_point_sampling(1, n, num_points, xyz.data_ptr(), temp.data_ptr(), idx.data_ptr(), _stream())
    return idx.long()


def to_ball_space(p: torch.Tensor):
    """WorldSpaceToBallSpace (:77-87) in float64 -> (normalised points, max_radius, center)"""
    center = (p.max(0).values + p.min(0).values) / 2
    d = p - center
    r = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]).sqrt().max()
    return d / r, r, center            # tensor / 0-dim device tensor: a true division


def relabel_map(present: Sequence[int]) -> Dict[int, int]:
    """The relabel loop of sample_and_save (:136-142) on the SET of instance labels that survived FPS: while j < max, a
    missing j takes over the current maximum.  -> {old label: new label} for the labels that move."""
    s = set(int(v) for v in present if v >= 0)
    cur = {v: v for v in s}            # current label -> original label
    j = 0
    while s and j < max(s):
        if j not in s:
            m = max(s)
            s.remove(m)
            s.add(j)
            cur[j] = cur.pop(m)
        j += 1
    return {orig: new for new, orig in cur.items() if orig != new}


def convert_labels(sem: torch.Tensor, ins: torch.Tensor):
    """sem + 1, ins -1 -> -100, instance ids made contiguous (:125-142).  One small device -> host read (the set of
    instance ids): this is an offline tool."""
    sem_c = sem + 1
    ins_c = ins.clone()
    ins_c[ins_c == -1] = -100
    present = torch.unique(ins_c[ins_c >= 0]).tolist()
    mp = relabel_map(present)
    if mp:
        hi = max(present) + 1
        lut = torch.arange(hi, device=ins.device, dtype=ins_c.dtype)
        for old, new in mp.items():
            lut[old] = new
        pos = ins_c >= 0
        ins_c[pos] = lut[ins_c[pos].long()]
    return sem_c, ins_c


def gt_labels(sem_c: torch.Tensor, ins_c: torch.Tensor) -> torch.Tensor:
    """evaluation labels (:158-169): semantic label OF THE INSTANCE'S FIRST POINT * 1000 + instance id, -100 elsewhere"""
    out = torch.full_like(ins_c, -100, dtype=torch.int32)
    pos = ins_c >= 0
    if not bool(pos.any()):
        return out
    n_inst = int(ins_c.max().item()) + 1
    first = torch.full((n_inst,), ins_c.numel(), dtype=torch.long, device=ins_c.device)
    first.scatter_reduce_(0, ins_c[pos].long(), pos.nonzero()[:, 0], reduce="amin")
    if bool((first == ins_c.numel()).any()):
        raise ValueError("a part is missing from the point cloud, instance label is not continuous")
    sem_of = sem_c[first]
    if bool((sem_of == 0).any()):
        raise ValueError("a part with semantic label [others]")
    out[pos] = (sem_of[ins_c[pos].long()] * MAX_INSTANCE_NUM + ins_c[pos]).int()
    return out


def sample_frame(pcs, rgb, sem, ins, npcs, idx, num_points: int) -> Optional[Dict[str, torch.Tensor]]:
    """sample_and_save (:112-156) minus the file writes: FPS, gather, ball normalisation, labels.  None if the frame
    has fewer than num_points points (the reference skips it, :113-114)."""
    if pcs.shape[0] < num_points:
        return None
    if ((sem == -1) != (ins == -1)).any():
        raise ValueError("Semantic and instance labels do not match!")
    f = fps(pcs, num_points)
    pn, r, center = to_ball_space(pcs[f])
    sem_c, ins_c = convert_labels(sem[f], ins[f])
    return dict(xyz=pn.float(), rgb=rgb[f].float(), sem=sem_c.int(), ins=ins_c.int(), npcs=npcs[f].float(), idx=idx[f].int(),
                scale_param=torch.cat([r.reshape(1), center]), gt=gt_labels(sem_c, ins_c), fps_idx=f)


# ---------------------------------------------------------------------------------------------------- shard format
def _align(v: int, a: int = 64) -> int:
    return (v + a - 1) // a * a


def _layout(S: int, N: int, name_bytes: int):
    off = 64
    lay = {}
    for name, dt, k in _SECTIONS:
        shape = (S, N, k) if k else (S, N)
        lay[name] = (off, np.dtype(dt), shape)
        off = _align(off + int(np.prod(shape)) * np.dtype(dt).itemsize)
    lay["scale"] = (off, np.dtype(np.float64), (S, 4))
    off = _align(off + S * 32)
    lay["names"] = (off, np.dtype(np.uint8), (S, name_bytes))
    off = _align(off + S * name_bytes)
    return lay, off


class ShardWriter:
    """with ShardWriter(path, num_frames, num_points) as w: w.add(pc_id, frame_dict) ..."""

    def __init__(self, path: str, num_frames: int, num_points: int, name_bytes: int = 64):
        self.path, self.S, self.N, self.nb = path, int(num_frames), int(num_points), int(name_bytes)
        lay, total = _layout(self.S, self.N, self.nb)
        with open(path, "wb") as f:
            f.truncate(total)
        self.mm = np.memmap(path, dtype=np.uint8, mode="r+")
        hdr = MAGIC + np.array([1, 0, self.N, self.nb, self.S], dtype="<u4").tobytes()
        self.mm[: len(hdr)] = np.frombuffer(hdr, dtype=np.uint8)
        self.views = {k: np.ndarray(shape, dtype=dt, buffer=self.mm, offset=off) for k, (off, dt, shape) in lay.items()}
        self.count = 0

    def add(self, pc_id: str, frame: Dict[str, torch.Tensor]) -> int:
        i = self.count
        if i >= self.S:
            raise GapartError("shard is full")
        for name, dt, _ in _SECTIONS:
            a = frame[name]
            a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
            if a.shape != self.views[name].shape[1:]:
                raise GapartError(f"{name}: shape {a.shape} does not fit the shard's {self.views[name].shape[1:]}")
            self.views[name][i] = a.astype(dt, copy=False)
        sp = frame["scale_param"]
        self.views["scale"][i] = sp.detach().cpu().numpy() if isinstance(sp, torch.Tensor) else np.asarray(sp)
        raw = pc_id.encode()
        if len(raw) >= self.nb:
            raise GapartError("pc_id too long for the shard's name field")
        self.views["names"][i] = 0
        self.views["names"][i, : len(raw)] = np.frombuffer(raw, dtype=np.uint8)
        self.count += 1
        return i

    def close(self):
        # the frame count goes in last: a shard whose writer died reads as empty; the sections keep their capacity layout
        self.mm[12:16] = np.frombuffer(np.array([self.count], dtype="<u4").tobytes(), dtype=np.uint8)
        self.mm.flush()
        del self.views, self.mm

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class ShardReader:
    def __init__(self, path: str):
        self.mm = np.memmap(path, dtype=np.uint8, mode="r")
        if bytes(self.mm[:8]) != MAGIC:
            raise GapartError(f"{path}: not a gapart shard")
        ver, count, N, nb, cap = (int(v) for v in np.frombuffer(bytes(self.mm[8:28]), dtype="<u4"))
        if ver != 1 or count > cap:
            raise GapartError(f"{path}: shard version {ver}, {count} frames of capacity {cap}")
        self.N, self.count = N, count
        lay, _ = _layout(cap, N, nb)
        self.views = {k: np.ndarray(shape, dtype=dt, buffer=self.mm, offset=off) for k, (off, dt, shape) in lay.items()}

    def __len__(self):
        return self.count

    def pc_id(self, i: int) -> str:
        return bytes(self.views["names"][i]).split(b"\0", 1)[0].decode()

    def pth_tuple(self, i: int):
        """the reference's .pth tuple (convert_rendered_into_input.py:156): zero-copy views into the mapped file"""
        if not 0 <= i < self.count:
            raise IndexError(i)
        return tuple(self.views[name][i] for name, _, _ in _SECTIONS)

    def scale_param(self, i: int) -> np.ndarray:
        return self.views["scale"][i]

    def load_data(self, i: int) -> Dict[str, object]:
        """the fields of the reference's load_data (dataset/gapartnet.py:208-229)"""
        xyz, rgb, sem, ins, npcs, _ = self.pth_tuple(i)
        return dict(pc_id=self.pc_id(i), points=np.concatenate([xyz, rgb], axis=-1, dtype=np.float32),
                    sem_labels=sem.astype(np.int64), instance_labels=ins.astype(np.int32), gt_npcs=npcs.astype(np.float32))
