"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the reference's
``GridSample`` voxel hash / unique (pointcept/datasets/transform.py:1181-1416).
Never imported by the product package.

Parity status: PINNED against the reference's own ``GridSample.__call__``
imported from /root/reference (tests/golden/gridsample_*.npz).

Implementation-independent outputs (bit-exact): voxel order (= ascending uint64
hash), ``inverse``, ``count``, ``grid_coord``.  Which member represents a voxel
in train mode depends on numpy's unstable introsort + the global RNG in the
reference (transform.py:1220,1264-1268); here -- and in the product kernels --
the rule is made explicit: members of a voxel are ordered by original index
(stable sort) and member ``r[v] % count[v]`` is taken, where ``r`` is the
``randint(0, count.max(), count.size)`` vector passed in by the caller.
"""
from __future__ import annotations

import numpy as np

FNV_OFFSET = np.uint64(14695981039346656037)
FNV_PRIME = np.uint64(1099511628211)


def voxelize(coord, grid_size: float):
    """transform.py:1213-1216.  NOTE the float64 divide (0-d float64 array operand
    under NumPy >= 2) -- an fp32 divide disagrees on ~8 voxel indices per million."""
    scaled = np.asarray(coord).astype(np.float64) / np.float64(grid_size)
    g = np.floor(scaled).astype(np.int64)
    gmin = g.min(0)
    return g - gmin, gmin


def fnv_hash_vec(arr) -> np.ndarray:
    """transform.py:1402-1416 (multiply THEN xor with the whole 64-bit coordinate)."""
    arr = np.asarray(arr).astype(np.uint64)
    h = np.full(arr.shape[0], FNV_OFFSET, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for j in range(arr.shape[1]):
            h = h * FNV_PRIME
            h = h ^ arr[:, j]
    return h


def ravel_hash_vec(arr) -> np.ndarray:
    """transform.py:1384-1399."""
    arr = np.asarray(arr).copy()
    arr -= arr.min(0)
    arr = arr.astype(np.uint64)
    arr_max = arr.max(0).astype(np.uint64) + np.uint64(1)
    keys = np.zeros(arr.shape[0], dtype=np.uint64)
    for j in range(arr.shape[1] - 1):
        keys += arr[:, j]
        keys *= arr_max[j + 1]
    keys += arr[:, -1]
    return keys


def grid_sample_index(coord, grid_size: float, hash_type: str = "fnv"):
    """Index part of GridSample.__call__ (transform.py:1211-1222).
    Returns dict(grid_coord_all[N,3], key[N] u64, idx_sort[N] (stable), inverse[N]
    (per raw point -> voxel rank in ascending-hash order, transform.py:1281-1283),
    count[M], start[M], min_coord[3])."""
    g, gmin = voxelize(coord, grid_size)
    key = fnv_hash_vec(g) if hash_type == "fnv" else ravel_hash_vec(g)
    idx_sort = np.argsort(key, kind="stable").astype(np.int64)
    key_sort = key[idx_sort]
    _, inv_sorted, count = np.unique(key_sort, return_inverse=True, return_counts=True)
    inverse = np.zeros_like(inv_sorted, dtype=np.int64)
    inverse[idx_sort] = inv_sorted
    start = np.cumsum(np.insert(count, 0, 0)[0:-1]).astype(np.int64)
    return dict(grid_coord_all=g, key=key, idx_sort=idx_sort, inverse=inverse,
                count=count.astype(np.int64), start=start, min_coord=gmin)


def grid_sample_train(coord, grid_size: float, rand=None, hash_type: str = "fnv"):
    """Train mode (transform.py:1264-1300).  ``rand`` = the reference's
    ``np.random.randint(0, count.max(), count.size)`` vector (None -> zeros, i.e. the
    lowest-index member).  Returns idx_unique[M], grid_coord[M,3], inverse[N], count[M]."""
    ix = grid_sample_index(coord, grid_size, hash_type)
    count = ix["count"]
    if rand is None:
        rand = np.zeros(count.size, dtype=np.int64)
    idx_select = ix["start"] + np.asarray(rand).astype(np.int64) % count
    idx_unique = ix["idx_sort"][idx_select]
    return dict(idx_unique=idx_unique, grid_coord=ix["grid_coord_all"][idx_unique],
                inverse=ix["inverse"], count=count, min_coord=ix["min_coord"])


def grid_sample_test(coord, grid_size: float, hash_type: str = "fnv"):
    """Test mode (transform.py:1302-1330): count.max() fragments, fragment i takes
    member ``i % count`` of every voxel.  Returns list of idx_part arrays + index dict."""
    ix = grid_sample_index(coord, grid_size, hash_type)
    parts = []
    for i in range(int(ix["count"].max())):
        idx_select = ix["start"] + i % ix["count"]
        parts.append(ix["idx_sort"][idx_select])
    return parts, ix
