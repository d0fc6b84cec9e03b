"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the reference's
serialization path.  Never imported by the product package.

Parity status: PINNED.  Every function here is checked bit-for-bit against the
reference's own code imported from /root/reference (see
tests/golden/make_golden.py, fixtures in tests/golden/serialization_*.npz).

Follows:
  * z-order       pointcept/models/utils/serialization/z_order.py:40-50,66-101
  * hilbert       pointcept/models/utils/serialization/hilbert.py:91-198 (Skilling)
  * encode        pointcept/models/utils/serialization/default.py:8-24
  * serialization pointcept/models/utils/structure.py:47-102
"""
from __future__ import annotations

import numpy as np

ORDERS = ("z", "z-trans", "hilbert", "hilbert-trans")


def z_order_key(x, y, z, depth: int) -> np.ndarray:
    """bit i of x -> 3i+2, y -> 3i+1, z -> 3i  (z_order.py:42-49); inputs masked to
    ``depth`` bits exactly like the LUT path does (z_order.py:90-95)."""
    x = np.asarray(x).astype(np.uint64)
    y = np.asarray(y).astype(np.uint64)
    z = np.asarray(z).astype(np.uint64)
    key = np.zeros_like(x)
    for i in range(depth):
        m = np.uint64(1 << i)
        key |= ((x & m) << np.uint64(2 * i + 2)) | ((y & m) << np.uint64(2 * i + 1)) | ((z & m) << np.uint64(2 * i))
    return key.astype(np.int64)


def hilbert_key(x, y, z, depth: int) -> np.ndarray:
    """Skilling transpose, MSB->LSB (hilbert.py:156-175), interleave with dim 0 most
    significant (hilbert.py:178), Gray->binary prefix XOR (hilbert.py:181,69-88)."""
    nb = depth
    mask = np.uint64((1 << nb) - 1)
    X = [np.asarray(v).astype(np.uint64) & mask for v in (x, y, z)]
    X = [v.copy() for v in X]
    q = 1 << (nb - 1)
    while q >= 1:
        Q = np.uint64(q)
        P = np.uint64(q - 1)
        for i in range(3):
            on = (X[i] & Q) != 0
            # bit on: invert low bits of dim 0
            inv = np.where(on, P, np.uint64(0))
            # bit off: exchange low bits of dim 0 and dim i
            t = np.where(on, np.uint64(0), (X[0] ^ X[i]) & P)
            X[0] = X[0] ^ inv ^ t
            if i != 0:
                X[i] = X[i] ^ t
        q >>= 1
    g = np.zeros_like(X[0])
    for b in range(nb):  # bit b of dim d -> position 3*b + (2-d)
        m = np.uint64(1 << b)
        for d in range(3):
            g |= ((X[d] & m) >> np.uint64(b)) << np.uint64(3 * b + (2 - d))
    h = g.copy()
    s = 1
    while s < 3 * nb:
        h ^= h >> np.uint64(s)
        s <<= 1
    return h.astype(np.int64)


def encode(grid_coord, batch=None, depth: int = 16, order: str = "z") -> np.ndarray:
    """serialization/default.py:8-24."""
    g = np.asarray(grid_coord).astype(np.int64)
    if order == "z":
        code = z_order_key(g[:, 0], g[:, 1], g[:, 2], depth)
    elif order == "z-trans":
        code = z_order_key(g[:, 1], g[:, 0], g[:, 2], depth)
    elif order == "hilbert":
        code = hilbert_key(g[:, 0], g[:, 1], g[:, 2], depth)
    elif order == "hilbert-trans":
        code = hilbert_key(g[:, 1], g[:, 0], g[:, 2], depth)
    else:
        raise NotImplementedError(order)
    if batch is not None:
        code = (np.asarray(batch).astype(np.int64) << np.int64(depth * 3)) | code
    return code


def offset2batch(offset) -> np.ndarray:
    """models/utils/misc.py:19-24."""
    offset = np.asarray(offset).astype(np.int64)
    counts = np.diff(np.concatenate([[0], offset]))
    return np.repeat(np.arange(len(offset), dtype=np.int64), counts)


def serialization_depth(grid_coord) -> int:
    """structure.py:66."""
    return int(np.asarray(grid_coord).max()).bit_length()


def serialization(grid_coord, batch, n_batch: int, order=ORDERS, depth=None, perm=None):
    """structure.py:47-102.  ``perm`` = the row permutation the reference draws with
    ``torch.randperm(len(order))`` when shuffle_orders=True (None = no shuffle).
    Ties (duplicate voxels) are broken by original index (stable), the
    deterministic rule the product kernels also use.
    Returns (code[k,N], order[k,N], inverse[k,N], depth)."""
    if depth is None:
        depth = serialization_depth(grid_coord)
    assert depth * 3 + int(n_batch).bit_length() <= 63
    assert depth <= 16
    code = np.stack([encode(grid_coord, batch, depth, o) for o in order])
    ordr = np.argsort(code, axis=1, kind="stable").astype(np.int64)
    inv = np.zeros_like(ordr)
    ar = np.arange(code.shape[1], dtype=np.int64)
    for k in range(code.shape[0]):
        inv[k, ordr[k]] = ar
    if perm is not None:
        perm = np.asarray(perm)
        code, ordr, inv = code[perm], ordr[perm], inv[perm]
    return code, ordr, inv, depth
