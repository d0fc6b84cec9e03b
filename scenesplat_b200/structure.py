"""`Point` -- host-side mirror of the reference's Point dict
(pointcept/models/utils/structure.py:14-140) on top of the sm_100a kernels.

Same keys, same attribute-and-item access, same method names (`serialization`, `sparsify`); the
work underneath is one fused encode + radix-sort launch sequence instead of ~hundreds of tiny torch
kernels, and `sparsify` no longer synchronises with the host (the spconv-style tensor computes
`indices` / `spatial_shape` / `batch_size` lazily, only if somebody reads them).
"""
from __future__ import annotations

import torch

from . import ops
from .spconv_compat import SparseConvTensor


class Dict(dict):
    """Attribute access == item access (the subset of addict.Dict the reference relies on)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        try:
            del self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def offset2bincount(offset):
    """pointcept/models/utils/misc.py:12-16"""
    return torch.diff(offset, prepend=torch.zeros(1, device=offset.device, dtype=torch.long))


def offset2batch(offset, n=None):
    """pointcept/models/utils/misc.py:19-24 (pass n to avoid the host sync of repeat_interleave)."""
    bincount = offset2bincount(offset)
    return torch.arange(len(bincount), device=offset.device, dtype=torch.long).repeat_interleave(
        bincount, output_size=n)


def batch2offset(batch):
    """pointcept/models/utils/misc.py:27-28"""
    return torch.cumsum(batch.bincount(), dim=0).long()


class Point(Dict):
    """See pointcept/models/utils/structure.py:14-45 for the key contract."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if "batch" not in self.keys() and "offset" in self.keys():
            n = None
            for k in ("coord", "grid_coord", "feat"):
                if k in self.keys():
                    n = self[k].shape[0]
                    break
            self["batch"] = offset2batch(self.offset, n)
        elif "offset" not in self.keys() and "batch" in self.keys():
            self["offset"] = batch2offset(self.batch)

    # ------------------------------------------------------------------ serialization
    def serialization(self, order="z", depth=None, shuffle_orders=False):
        """structure.py:47-102.  `torch.randperm` is drawn from the default CPU generator exactly like
        the reference, so a seeded run consumes the RNG stream identically."""
        assert "batch" in self.keys()
        if "grid_coord" not in self.keys():
            assert {"grid_size", "coord"}.issubset(self.keys())
            self["grid_coord"] = torch.div(
                self.coord - self.coord.min(0)[0], self.grid_size, rounding_mode="trunc").int()
        order = [order] if isinstance(order, str) else list(order)
        if depth is None:
            depth = ops.coord_depth(self.grid_coord)
        self["serialized_depth"] = depth
        assert depth * 3 + len(self.offset).bit_length() <= 63
        assert depth <= 16
        if shuffle_orders:
            perm = torch.randperm(len(order)).tolist()
            order = [order[i] for i in perm]
        _, code, ordr, inverse = ops.serialize(self.grid_coord, self.offset, depth, order, want_batch=False)
        self["serialized_code"] = code
        self["serialized_order"] = ordr
        self["serialized_inverse"] = inverse
        self["serialized_order_names"] = tuple(order)  # row r was encoded with this curve (kernel-map lookups)

    # ------------------------------------------------------------------ sparsify
    def sparsify(self, pad=96):
        """structure.py:104-140, without the two host syncs."""
        assert {"feat", "batch"}.issubset(self.keys())
        if "grid_coord" not in self.keys():
            assert {"grid_size", "coord"}.issubset(self.keys())
            self["grid_coord"] = torch.div(
                self.coord - self.coord.min(0)[0], self.grid_size, rounding_mode="trunc").int()
        self["sparse_conv_feat"] = SparseConvTensor(
            features=self.feat, indices=None, spatial_shape=self.get("sparse_shape", None), batch_size=None,
            point=self, pad=pad)

    def octreetization(self, depth=None, full_depth=None):
        raise NotImplementedError("octree path needs ocnn and is unused by PT-v3m1 (out of scope, SURVEY.md 2.1)")
