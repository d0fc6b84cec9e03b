"""Call-compatible stand-ins for the three spconv symbols the reference's PTv3 touches
(`spconv.pytorch.SubMConv3d`, `SparseConvTensor`, `spconv.pytorch.modules.is_spconv_module`;
call sites pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py:277-284,499-506,
pointcept/models/utils/structure.py:131-138, pointcept/models/modules.py:64-75), backed by the
kernel-map + tcgen05 gather-GEMM kernels.  Parameter names / shapes match spconv >= 2.2
(`weight [Cout, k, k, k, Cin]`, `bias [Cout]`) so reference checkpoints load unchanged.
"""
from __future__ import annotations

import math
import weakref

import torch
import torch.nn as nn

from . import ops


class SparseConvTensor:
    """features + lazily materialised indices / spatial_shape / batch_size (no host sync unless read)."""

    def __init__(self, features, indices, spatial_shape, batch_size, point=None, pad=96, indice_dict=None):
        self.features = features
        self._indices = indices
        self._spatial_shape = spatial_shape
        self._batch_size = batch_size
        # The owning Point is held WEAKLY: the Point holds this tensor (`sparse_conv_feat`), and a strong reference back
        # would make every Point of every forward cyclic garbage that only the cyclic collector frees -- with GBs of
        # device tensors attached (measured: the caching allocator grew by 2.4 GB per step until the next full collection).
        self._point_ref = weakref.ref(point) if point is not None else None
        self._own_point = None  # private z-order view of a tensor built without a Point (point_view)
        self._pad = pad
        self.indice_dict = indice_dict if indice_dict is not None else {}

    @property
    def _point(self):
        if self._own_point is not None:
            return self._own_point
        p = self._point_ref() if self._point_ref is not None else None
        if p is None and self._point_ref is not None:
            raise RuntimeError("SparseConvTensor: the Point this tensor was built from no longer exists")
        return p

    @property
    def indices(self):
        if self._indices is None:
            p = self._point
            self._indices = torch.cat([p.batch.unsqueeze(-1).int(), p.grid_coord.int()], dim=1).contiguous()
        return self._indices

    @property
    def spatial_shape(self):
        if self._spatial_shape is None:
            self._spatial_shape = torch.add(torch.max(self._point.grid_coord, dim=0).values, self._pad).tolist()
            self._point["sparse_shape"] = self._spatial_shape
        return self._spatial_shape

    @property
    def batch_size(self):
        if self._batch_size is None:
            self._batch_size = int(self._point.batch[-1]) + 1
        return self._batch_size

    def point_view(self):
        """The Point whose serialized codes index this tensor's voxels.  A tensor built the spconv way
        (features + [N, 4] (batch, z, y, x) indices, no Point; structure.py:131-138) gets a private z-order view."""
        if self._point is None:
            from .structure import Point
            idx = self._indices
            p = Point(grid_coord=idx[:, 1:].contiguous(), batch=idx[:, 0].long().contiguous())
            p.serialization(order=("z",))
            self._own_point = p
        return self._point

    def replace_feature(self, feat):
        t = SparseConvTensor(feat, self._indices, self._spatial_shape, self._batch_size, None, self._pad, self.indice_dict)
        t._point_ref, t._own_point = self._point_ref, self._own_point
        return t


def kernel_map_for(point, k: int, want_pairs: bool):
    """Kernel map of a Point level, cached on the Point (the reference caches indice pairs per
    `indice_key` on the tensor lineage; encoder stage s and decoder stage s share the Point, hence the map)."""
    cache = point.setdefault("_kmap_cache", {})
    ent = cache.get(k)
    if ent is None:
        names = point.serialized_order_names
        row = next((i for i, nm in enumerate(names) if nm in ("z", "z-trans")), 0)
        k_big = next((kk for kk in sorted(cache) if kk > k and (kk - k) % 2 == 0 and cache[kk]["row"] == row), None)
        if k_big is not None:  # a larger map of the same voxels exists (the stem's 5^3): its rows contain this window
            nbr, cnt = ops.kmap_subset(cache[k_big]["nbr"], cache[k_big]["count"], k_big, k)
        else:
            nbr, cnt = _kmap_search(point, row, k)
        ent = dict(nbr=nbr, count=cnt, row=row, pairs=None)
        cache[k] = ent
    if want_pairs and ent["pairs"] is None:
        ent["pairs"] = ops.kmap_pairs(ent["nbr"], point.serialized_order[ent["row"]], k, ent["count"].cpu().numpy())
    return ent


def _kmap_search(point, row, k):
    names = point.serialized_order_names
    return ops.kmap_build(point.grid_coord, point.batch, point.serialized_code[row], point.serialized_order[row],
                          point.serialized_depth, ops.ORDER_IDS[names[row]], k)


class SubMConv3d(nn.Module):
    """Submanifold conv; `padding` is accepted and ignored like spconv does for SubM convs."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, **kw):
        super().__init__()
        assert stride == 1 and dilation == 1 and groups == 1
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.indice_key = indice_key
        k = kernel_size
        self.weight = nn.Parameter(torch.empty(out_channels, k, k, k, in_channels))
        nn.init.kaiming_uniform_(self.weight.view(out_channels, -1), a=math.sqrt(5))
        if bias:
            bound = 1 / math.sqrt(in_channels * k ** 3)
            self.bias = nn.Parameter(torch.empty(out_channels).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)
        self._cache = {}

    def _prepared(self, kind):
        ver = (self.weight._version, self.weight.data_ptr(), self.weight.device)
        ent = self._cache.get(kind)
        if ent is None or ent[0] != ver:
            k3 = self.kernel_size ** 3
            w = self.weight.detach().reshape(self.out_channels, k3, self.in_channels)
            if kind == "tc":  # [k3, cout, cin] bf16, K-major rows for TMA
                t = w.permute(1, 0, 2).contiguous().to(torch.bfloat16)
            else:  # [k3, cin, cout] fp32 for the SIMT kernel
                t = w.permute(1, 2, 0).contiguous().float()
            ent = (ver, t)
            self._cache[kind] = ent
        return ent[1]

    def tensor_core_ok(self):
        return self.in_channels % 16 == 0 and self.out_channels % 32 == 0 and self.in_channels >= 32

    def conv_point(self, point, feat, scale=None, shift=None, act=0, out_dtype=None):
        """Convolve `feat` (rows = the Point's voxels).  Optional fused folded-BN affine + activation
        (SIMT path only; used by the stem)."""
        if torch.is_grad_enabled() and (feat.requires_grad or self.weight.requires_grad):
            return self._conv_autograd(point, feat, scale, act, out_dtype)
        bias = self.bias.detach().float() if self.bias is not None else None
        if self.tensor_core_ok() and scale is None and act == 0:
            ent = kernel_map_for(point, self.kernel_size, want_pairs=True)
            x = feat if feat.dtype == torch.bfloat16 else feat.to(torch.bfloat16)
            return ops.subm_conv_gemm(x, ent["pairs"], self._prepared("tc"), bias, feat.shape[0],
                                      out_dtype=out_dtype or torch.bfloat16)
        ent = kernel_map_for(point, self.kernel_size, want_pairs=False)
        x = feat if feat.dtype in (torch.float32, torch.bfloat16) else feat.float()
        return ops.subm_conv_simt(x, ent["nbr"], self._prepared("simt"), bias, scale, shift, act,
                                  out_dtype=out_dtype or torch.float32)

    def _conv_autograd(self, point, feat, scale, act, out_dtype):
        """Differentiable path (scenesplat_b200/training.py): tensor-core conv with dgrad / wgrad kernels, or the stem
        form (tiny Cin, input without gradient)."""
        from . import training as T
        if scale is not None or act != 0:
            raise NotImplementedError("fused affine / activation epilogues are inference-only")
        if self.tensor_core_ok():
            ent = kernel_map_for(point, self.kernel_size, want_pairs=True)
            y = T.SubMConvFn.apply(feat.to(torch.bfloat16), self.weight, self.bias, ent["pairs"], feat.shape[0])
        elif not feat.requires_grad:
            ent = kernel_map_for(point, self.kernel_size, want_pairs=False)
            y = T.StemConvFn.apply(feat.float(), self.weight, ent["nbr"])
            if self.bias is not None:
                y = y + self.bias
        else:
            raise NotImplementedError("SubMConv3d input gradients need Cin % 16 == 0, Cin >= 32 and Cout % 32 == 0")
        return y.to(out_dtype) if out_dtype is not None else y

    def forward(self, x: SparseConvTensor):
        return x.replace_feature(self.conv_point(x.point_view(), x.features, out_dtype=torch.float32))


def is_spconv_module(module):
    return isinstance(module, SubMConv3d)
