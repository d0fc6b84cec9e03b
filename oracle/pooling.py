"""TEST INFRASTRUCTURE ONLY -- CPU restatement of SerializedPooling /
SerializedUnpooling index building and segmented reduction
(pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py:371-444,471-482).
Never imported by the product package.

Parity status: index outputs PINNED against the reference's own
``SerializedPooling.forward`` (tests/golden/pooling_*.npz).  The segmented mean
itself lives in ``torch_scatter.segment_csr`` which is absent from the reference
tree and from this image -> its arithmetic (accumulation dtype under AMP) is
"parity unpinned"; this oracle accumulates in fp32 over rows
``indptr[i]:indptr[i+1]`` as the published semantics of segment_csr state.
"""
from __future__ import annotations

import numpy as np


def pool_index(code, pooling_depth: int = 1, perm=None):
    """ptv3:384-412.  ``code`` [k,N] int64 (row 0 decides clusters).
    Returns dict(cluster[N], counts[M], indices[N] (points sorted by cluster, stable),
    idx_ptr[M+1], head[M], code[k,M], order[k,M], inverse[k,M])."""
    code = np.asarray(code).astype(np.int64) >> np.int64(pooling_depth * 3)
    _, cluster, counts = np.unique(code[0], return_inverse=True, return_counts=True)
    cluster = cluster.astype(np.int64)
    indices = np.argsort(cluster, kind="stable").astype(np.int64)
    idx_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    head = indices[idx_ptr[:-1]]
    pcode = code[:, head]
    order = np.argsort(pcode, axis=1, kind="stable").astype(np.int64)
    inverse = np.zeros_like(order)
    ar = np.arange(pcode.shape[1], dtype=np.int64)
    for k in range(pcode.shape[0]):
        inverse[k, order[k]] = ar
    if perm is not None:
        perm = np.asarray(perm)
        pcode, order, inverse = pcode[perm], order[perm], inverse[perm]
    return dict(cluster=cluster, counts=counts.astype(np.int64), indices=indices, idx_ptr=idx_ptr,
                head=head, code=pcode, order=order, inverse=inverse)


def segment_csr(src, indices, idx_ptr, reduce: str = "mean"):
    """torch_scatter.segment_csr(src[indices], idx_ptr, reduce) with fp32 accumulation in
    member order (ptv3:416-421)."""
    src = np.asarray(src, dtype=np.float32)[indices]
    m = len(idx_ptr) - 1
    out = np.zeros((m,) + src.shape[1:], dtype=np.float32)
    if reduce in ("sum", "mean"):
        np.add.at(out, np.repeat(np.arange(m), np.diff(idx_ptr)), src)
        if reduce == "mean":
            out /= np.diff(idx_ptr).astype(np.float32).reshape((-1,) + (1,) * (src.ndim - 1))
    elif reduce == "max":
        out[:] = -np.inf
        np.maximum.at(out, np.repeat(np.arange(m), np.diff(idx_ptr)), src)
    elif reduce == "min":
        out[:] = np.inf
        np.minimum.at(out, np.repeat(np.arange(m), np.diff(idx_ptr)), src)
    else:
        raise ValueError(reduce)
    return out


def pooled_attrs(grid_coord, batch, head, pooling_depth: int = 1):
    """ptv3:422,427."""
    return (np.asarray(grid_coord)[head] >> pooling_depth), np.asarray(batch)[head]


def unpool_gather_add(parent_feat, child_feat, cluster):
    """ptv3:478: parent.feat + point.feat[pooling_inverse]."""
    return np.asarray(parent_feat) + np.asarray(child_feat)[cluster]
