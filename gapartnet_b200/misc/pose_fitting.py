"""NPCS -> part pose (similarity transform + oriented bbox), CPU/numpy like the reference
(/root/reference/gapartnet/misc/pose_fitting.py: Umeyama :4-39, RANSAC :54-80, transform :83-118,
estimate_pose_from_npcs :121-147).  BASELINE config #1 plumbing; it consumes numpy's global RNG in
the same order as the reference (one `randint(n, size=5)` per RANSAC iteration) so that, given the
same seed, the results are identical (tests/golden/pose_cfg1.npz was produced by the reference)."""
from __future__ import annotations

import numpy as np


def umeyama(src: np.ndarray, dst: np.ndarray):
    """least-squares similarity dst ~ s*R*src + t for [3,n] arrays (reflection-corrected SVD).
    Returns (scale[3], rotation (row-vector convention, as the reference), translation, T[4,4])."""
    n = src.shape[1]
    mu_s, mu_d = src.mean(axis=1), dst.mean(axis=1)
    cs, cd = src - mu_s[:, None], dst - mu_d[:, None]
    cov = cd @ cs.T / n
    if np.isnan(cov).any():
        raise RuntimeError("There are NANs in the input.")
    U, D, Vh = np.linalg.svd(cov, full_matrices=True)
    if np.linalg.det(U) * np.linalg.det(Vh) < 0.0:
        D[-1] = -D[-1]
        U[:, -1] = -U[:, -1]
    s = D.sum() / np.var(src, axis=1).sum()
    R = (U @ Vh).T
    t = mu_d - mu_s.dot(s * R)
    T = np.identity(4)
    T[:3, :3] = np.diag([s, s, s]) @ R
    T[:3, 3] = t
    return np.array([s, s, s]), R, t, T


def _residuals(T, src_h, dst_h):
    return np.linalg.norm((dst_h - T @ src_h)[:3], axis=0)


def ransac_inliers(src_h, dst_h, max_iters: int, pass_thrsh: float, stop_thrsh: float):
    n = src_h.shape[1]
    best_res, best_ratio, best_idx = 1e10, 0, np.arange(n)
    for _ in range(max_iters):
        pick = np.random.randint(n, size=5)
        _, _, _, T = umeyama(src_h[:3, pick], dst_h[:3, pick])
        r = _residuals(T, src_h, dst_h)
        res = np.linalg.norm(r)
        idx = np.where(r < pass_thrsh)[0]
        # the reference counts non-zero inlier *indices* (index 0 never counts), misc/pose_fitting.py:49
        ratio = np.count_nonzero(idx) / n
        if res < best_res:
            best_res, best_ratio, best_idx = res, ratio, idx
        if best_res < stop_thrsh:
            break
    return best_ratio, best_idx


def estimate_similarity_transform(source, target, stop_thrsh: float = 0.5, max_iters: int = 100):
    if source.shape[0] == 1:
        source, target = np.repeat(source, 2, axis=0), np.repeat(target, 2, axis=0)
    src_h = np.vstack([source.T, np.ones(source.shape[0])])
    dst_h = np.vstack([target.T, np.ones(target.shape[0])])
    ns = np.mean(np.linalg.norm(source, axis=1))
    nt = np.mean(np.linalg.norm(target, axis=1))
    pass_thrsh = max(ns / nt, nt / ns)
    ratio, idx = ransac_inliers(src_h, dst_h, max_iters, pass_thrsh, stop_thrsh)
    if ratio < 0.01:
        return np.asarray([None, None, None]), None, None, None, None
    s, R, t, T = umeyama(src_h[:3, idx], dst_h[:3, idx])
    return s, R, t, T, idx


def estimate_pose_from_npcs(xyz, npcs):
    """-> (bbox [8,3], scale[3], rotation, translation, T[4,4], inlier idx); bbox = inlier extent in
    canonical space mapped back to the camera frame"""
    s, R, t, T, idx = estimate_similarity_transform(npcs, xyz)
    if s[0] is None:
        return None, np.asarray([None, None, None]), None, None, None, idx
    canon = np.dot(xyz - t, np.linalg.pinv(R)) / s[0]
    ext = np.abs(canon[idx]).max(0)
    signs = np.array([[-1, -1, -1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1], [1, 1, -1], [1, -1, 1], [-1, 1, 1], [1, 1, 1]])
    bbox = np.dot(signs * ext * s[0], R) + t
    return bbox, s, R, t, T, idx
