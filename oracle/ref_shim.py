"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference sources.

Imports SceneSplat's own Python files from ``/root/reference`` (read-only,
present in the build container only, never on the GPU box) so that

  * ``tests/golden/make_golden.py`` can generate golden vectors from the
    reference itself, and
  * ``tests/test_oracle_vs_reference.py`` can pin ``oracle/`` against it.

The reference's hot path needs four packages that are not installed in this
image (``addict``, ``timm``, ``torch_scatter``, ``spconv``).  They are provided
as small stand-ins below, exactly following the published semantics of the
call sites:

  * ``torch_scatter.segment_csr(src, indptr, reduce)``: row segment ``i``
    reduces ``src[indptr[i]:indptr[i+1]]``
    (call sites ``point_transformer_v3m1_base.py:416-421``).
  * ``spconv.pytorch.SubMConv3d``: submanifold conv, cross-correlation,
    weight ``[Cout, k, k, k, Cin]`` (spconv >= 2.2 KRSC layout), output sites
    = input sites (call sites ``point_transformer_v3m1_base.py:277-284,499-506``).
    The arithmetic of the real library cannot be exercised here -> everything
    downstream of this shim is "parity unpinned" w.r.t. real spconv.
  * ``addict.Dict``, ``timm.layers.DropPath``: trivial.

Nothing in the product package imports this module.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("SCENESPLAT_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pointcept"))


# --------------------------------------------------------------------------- addict
class _Dict(dict):
    """Minimal addict.Dict: attribute access == item access."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for a in args:
            if not a:
                continue
            if isinstance(a, dict):
                for k, v in a.items():
                    self[k] = v
            else:
                for k, v in a:
                    self[k] = v
        for k, v in kwargs.items():
            self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # addict returns an empty Dict; the path never relies on it
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]


# --------------------------------------------------------------------------- timm
class _DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


# --------------------------------------------------------------------------- torch_scatter
def _segment_csr(src, indptr, out=None, reduce="sum"):
    lengths = torch.diff(indptr)
    red = {"sum": "sum", "mean": "mean", "min": "min", "max": "max"}[reduce]
    return torch.segment_reduce(src, red, lengths=lengths, axis=0, unsafe=True)


# --------------------------------------------------------------------------- spconv
class _SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, indice_dict=None):
        self.features = features
        self.indices = indices
        self.spatial_shape = spatial_shape
        self.batch_size = batch_size
        self.indice_dict = indice_dict if indice_dict is not None else {}

    def replace_feature(self, feat):
        t = _SparseConvTensor(feat, self.indices, self.spatial_shape, self.batch_size, self.indice_dict)
        return t


def _kernel_map_dense(indices: torch.Tensor, k: int):
    """[N, k^3] int64 neighbour table (-1 = inactive) from (b,x,y,z) int indices."""
    idx = indices.long()
    n = idx.shape[0]
    r = k // 2
    mx = idx[:, 1:].max(0).values + 1 + 2 * r
    sx, sy, sz = (int(v) for v in mx)

    def lin(b, x, y, z):
        return ((b * sx + (x + r)) * sy + (y + r)) * sz + (z + r)

    base = lin(idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3])
    order = torch.argsort(base)
    skeys = base[order]
    nbr = torch.full((n, k * k * k), -1, dtype=torch.long)
    t = 0
    for dx in range(-r, r + 1):
        for dy in range(-r, r + 1):
            for dz in range(-r, r + 1):
                q = lin(idx[:, 0], idx[:, 1] + dx, idx[:, 2] + dy, idx[:, 3] + dz)
                pos = torch.searchsorted(skeys, q).clamp(max=n - 1)
                hit = skeys[pos] == q
                nbr[:, t] = torch.where(hit, order[pos], torch.full_like(pos, -1))
                t += 1
    return nbr


class _SubMConv3d(nn.Module):
    """Submanifold conv stand-in: out[p] = b + sum_t W[:, t, :] @ in[p + delta_t]."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=0,
                 dilation=1, groups=1, bias=True, indice_key=None, **kw):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.indice_key = indice_key
        k = kernel_size
        self.weight = nn.Parameter(torch.empty(out_channels, k, k, k, in_channels))
        nn.init.kaiming_uniform_(self.weight.view(out_channels, -1), a=5 ** 0.5)
        if bias:
            fan_in = in_channels * k ** 3
            bound = 1 / fan_in ** 0.5
            self.bias = nn.Parameter(torch.empty(out_channels).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)

    def forward(self, x: _SparseConvTensor):
        key = (self.indice_key, self.kernel_size)
        if self.indice_key is not None and key in x.indice_dict:
            nbr = x.indice_dict[key]
        else:
            nbr = _kernel_map_dense(x.indices, self.kernel_size)
            if self.indice_key is not None:
                x.indice_dict[key] = nbr
        feat = x.features
        k3 = self.kernel_size ** 3
        w = self.weight.reshape(self.out_channels, k3, self.in_channels)
        out = feat.new_zeros(feat.shape[0], self.out_channels)
        # spconv is a custom op: autocast does not touch it (the SSL variant only runs under AMP)
        with torch.autocast(device_type=feat.device.type, enabled=False):
            for t in range(k3):
                col = nbr[:, t]
                rows = (col >= 0).nonzero(as_tuple=True)[0]
                if rows.numel() == 0:
                    continue
                out.index_add_(0, rows, feat[col[rows]] @ w[:, t, :].t().to(feat.dtype))
        if self.bias is not None:
            out = out + self.bias
        return x.replace_feature(out)


def _install_shims():
    if "addict" not in sys.modules:
        m = types.ModuleType("addict")
        m.Dict = _Dict
        sys.modules["addict"] = m
    if "timm" not in sys.modules:
        m = types.ModuleType("timm")
        ml = types.ModuleType("timm.layers")
        ml.DropPath = _DropPath
        m.layers = ml
        sys.modules["timm"] = m
        sys.modules["timm.layers"] = ml
        mm_ = types.ModuleType("timm.models")  # the SSL variant imports DropPath from timm.models.layers
        mm_.layers = ml
        m.models = mm_
        sys.modules["timm.models"] = mm_
        sys.modules["timm.models.layers"] = ml
    if "torch_scatter" not in sys.modules:
        m = types.ModuleType("torch_scatter")
        m.segment_csr = _segment_csr
        sys.modules["torch_scatter"] = m
    if "spconv" not in sys.modules:
        m = types.ModuleType("spconv")
        mp = types.ModuleType("spconv.pytorch")
        mm = types.ModuleType("spconv.pytorch.modules")
        mm.is_spconv_module = lambda mod: isinstance(mod, _SubMConv3d)
        mp.SubMConv3d = _SubMConv3d
        mp.SparseConvTensor = _SparseConvTensor
        mp.modules = mm
        m.pytorch = mp
        sys.modules["spconv"] = m
        sys.modules["spconv.pytorch"] = mp
        sys.modules["spconv.pytorch.modules"] = mm
    for name, attrs in (("SharedArray", {}), ("termcolor", {"colored": lambda s, *a, **k: s})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m


_LOADED = {}


def load_reference():
    """Return a namespace with the reference's hot-path symbols (imported unmodified)."""
    if _LOADED:
        return types.SimpleNamespace(**_LOADED)
    if not reference_available():
        raise RuntimeError("reference tree not available at %s" % REFERENCE_ROOT)
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # pre-seed package stubs whose real __init__ import every backbone
    for pkg, rel in (("pointcept.models", "pointcept/models"),
                     ("pointcept.models.point_prompt_training", "pointcept/models/point_prompt_training"),
                     ("pointcept.models.losses", "pointcept/models/losses"),
                     ("pointcept.models.point_transformer_v3", "pointcept/models/point_transformer_v3")):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
            sys.modules[pkg] = m
    import pointcept  # noqa: F401  (real package __init__ is empty)
    pdn = importlib.import_module("pointcept.models.point_prompt_training.prompt_driven_normalization")
    sys.modules["pointcept.models.point_prompt_training"].PDNorm = pdn.PDNorm
    ser = importlib.import_module("pointcept.models.utils.serialization")
    structure = importlib.import_module("pointcept.models.utils.structure")
    ptv3 = importlib.import_module("pointcept.models.point_transformer_v3.point_transformer_v3m1_base")
    lb = importlib.import_module("pointcept.models.losses.builder")
    losses = importlib.import_module("pointcept.models.losses.misc")
    transform = importlib.import_module("pointcept.datasets.transform")
    dutils = importlib.import_module("pointcept.datasets.utils")
    _LOADED.update(
        encode=ser.encode, Point=structure.Point, ptv3=ptv3,
        PointTransformerV3=ptv3.PointTransformerV3, SerializedPooling=ptv3.SerializedPooling,
        SerializedAttention=ptv3.SerializedAttention, Block=ptv3.Block,
        losses=losses, Criteria=lb.Criteria, GridSample=transform.GridSample, SphereCrop=transform.SphereCrop,
        collate_fn=dutils.collate_fn, SubMConv3d=_SubMConv3d, SparseConvTensor=_SparseConvTensor,
        kernel_map_dense=_kernel_map_dense,
    )
    return types.SimpleNamespace(**_LOADED)


def load_reference_ssl():
    """The reference's self-supervised PT-v3m1 (point_transformer_v3_ssl/point_transformer_v3m1_ssl.py), unmodified."""
    load_reference()
    pkg = "pointcept.models.point_transformer_v3_ssl"
    if pkg not in sys.modules:
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REFERENCE_ROOT, "pointcept/models/point_transformer_v3_ssl")]
        sys.modules[pkg] = m
    mod = importlib.import_module(pkg + ".point_transformer_v3m1_ssl")
    return mod.PointTransformerV3_SIMDINO


LANG_BACKBONE_CFG = dict(
    in_channels=11,
    order=("z", "z-trans", "hilbert", "hilbert-trans"),
    stride=(2, 2, 2),
    enc_depths=(2, 2, 2, 6),
    enc_channels=(32, 64, 128, 256),
    enc_num_head=(2, 4, 8, 16),
    enc_patch_size=(1024, 1024, 1024, 1024),
    dec_depths=(2, 2, 2),
    dec_channels=(768, 512, 256),
    dec_num_head=(16, 16, 16),
    dec_patch_size=(1024, 1024, 1024),
    mlp_ratio=4, qkv_bias=True, qk_scale=None, attn_drop=0.0, proj_drop=0.0,
    drop_path=0.3, shuffle_orders=True, pre_norm=True, enable_rpe=False,
    enable_flash=True, upcast_attention=False, upcast_softmax=False, cls_mode=False,
)
