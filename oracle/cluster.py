"""CPU oracle (test infrastructure) - epic_ops clustering / scoring ops, restated from their call
sites (epic_ops itself is not vendored: parity unpinned, see oracle/__init__.py):

  ball_query      /root/reference/gapartnet/network/grouping_utils.py:119-128 (and the linear-scan
                  semantics of the reference's own pointnet2 ball_query_gpu.cu:9-45: strict `<`,
                  ascending point index, stop at the cap)
  ccl             grouping_utils.py:131-137: components of the (begin,end)-addressed adjacency table
  segmented_*     grouping_utils.py:59-70, network/model.py:360-362
  instance_iou    network/model.py:373-378
  nms             grouping_utils.py:244 (greedy, threshold on a dense IoU matrix)
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components


def ball_query(points, query, batch_indices, batch_offsets, radius, num_samples, point_labels=None,
               query_labels=None):
    pts = np.asarray(points, dtype=np.float32)
    qry = np.asarray(query, dtype=np.float32)
    Q = qry.shape[0]
    idx = np.full((Q, num_samples), -1, dtype=np.int32)
    num = np.zeros(Q, dtype=np.int32)
    r2 = np.float32(radius) * np.float32(radius)
    for q in range(Q):
        b = int(batch_indices[q])
        s, e = int(batch_offsets[b]), int(batch_offsets[b + 1])
        d = qry[q][None, :] - pts[s:e]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]   # same fp32 evaluation order
        ok = d2 < r2
        if point_labels is not None:
            ok &= np.asarray(point_labels[s:e]) == query_labels[q]
        hits = np.nonzero(ok)[0][:num_samples] + s
        idx[q, : hits.shape[0]] = hits
        num[q] = hits.shape[0]
    return idx, num


def ccl(offsets_flat, edges_flat):
    """labels[v] = smallest vertex index of v's component (edges undirected)."""
    off = np.asarray(offsets_flat).reshape(-1, 2)
    edges = np.asarray(edges_flat)
    V = off.shape[0]
    rows, cols = [], []
    for v in range(V):
        e = edges[off[v, 0]:off[v, 1]]
        e = e[(e >= 0) & (e < V)]
        rows.append(np.full(e.shape[0], v))
        cols.append(e)
    rows = np.concatenate(rows) if rows else np.zeros(0, int)
    cols = np.concatenate(cols) if cols else np.zeros(0, int)
    g = coo_matrix((np.ones(rows.shape[0]), (rows, cols)), shape=(V, V))
    _, comp = connected_components(g, directed=False)
    mins = np.full(comp.max() + 1 if V else 0, V, dtype=np.int64)
    np.minimum.at(mins, comp, np.arange(V))
    return mins[comp].astype(np.int32)


def segmented_reduce(x, begin, end, mode):
    x = np.asarray(x, dtype=np.float32)
    out = np.zeros((len(begin), x.shape[1]), dtype=np.float32)
    arg = np.full((len(begin), x.shape[1]), -1, dtype=np.int32)
    for s, (b, e) in enumerate(zip(begin, end)):
        if e <= b:
            continue
        seg = x[b:e]
        if mode == "sum":
            # fp64 accumulation, rounded to fp32 once: independent of the summation order (the kernel
            # reduces a segment in parallel the same way)
            out[s] = seg.astype(np.float64).sum(0).astype(np.float32)
        elif mode == "min":
            out[s] = seg.min(0)
        else:
            out[s] = seg.max(0)
            arg[s] = seg.argmax(0) + b          # first maximum
    return out, arg


def instance_iou(proposal_offsets, instance_labels, batch_indices, num_points_per_instance):
    po = np.asarray(proposal_offsets)
    npi = np.asarray(num_points_per_instance)
    P, Imax = po.shape[0] - 1, npi.shape[1]
    out = np.zeros((P, Imax), dtype=np.float32)
    for p in range(P):
        b0, e0 = po[p], po[p + 1]
        if e0 <= b0:
            continue
        lab = np.asarray(instance_labels[b0:e0])
        cnt = np.bincount(lab[(lab >= 0) & (lab < Imax)], minlength=Imax)
        n = npi[int(batch_indices[b0])]
        uni = (e0 - b0) + n - cnt
        out[p] = np.where(uni > 0, cnt.astype(np.float32) / np.maximum(uni, 1).astype(np.float32), 0)
    return out


def nms(ious, scores, threshold):
    order = np.argsort(-np.asarray(scores), kind="stable")
    dead = np.zeros(len(order), dtype=bool)
    keep = []
    for i, a in enumerate(order):
        if dead[i]:
            continue
        keep.append(a)
        dead[i + 1:] |= np.asarray(ious)[a, order[i + 1:]] > threshold
    return np.asarray(keep, dtype=np.int64)


def cluster_proposals(pt_xyz, batch_indices, batch_offsets, sem_preds, radius, cap):
    """cluster_proposals (grouping_utils.py:108-140): -> (sorted_cc_labels, sorted_indices) with a
    stable sort (ascending point index inside a label)."""
    idx, num = ball_query(pt_xyz, pt_xyz, batch_indices, batch_offsets, radius, cap, sem_preds, sem_preds)
    Q = idx.shape[0]
    begin = np.arange(Q, dtype=np.int64) * cap
    off = np.stack([begin, begin + num], axis=1).reshape(-1)
    labels = ccl(off, idx.reshape(-1))
    order = np.argsort(labels, kind="stable")
    return labels[order], order
