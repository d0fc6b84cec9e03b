"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the submanifold sparse
convolution the reference obtains from ``spconv.pytorch.SubMConv3d``
(call sites pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py:277-284
(xCPE 3^3, bias) and :499-506 (stem 5^3, no bias); tensor built at
pointcept/models/utils/structure.py:131-138; fp32 forced at pointcept/models/modules.py:68-74).
Never imported by the product package.

Parity status: UNPINNED w.r.t. real spconv -- the library (``spconv-cu124``,
unpinned, ``env.yaml:49``) is third party, absent from /root/reference and from this
image, and the reference holds no test for it.  Restated from spconv's published
semantics: output sites = input sites; ``out[p] = bias + sum_t W[:, t, :] @ in[p + delta_t]``
over active neighbours; cross-correlation (no kernel flip); weight layout
``[Cout, k, k, k, Cin]`` with the three kernel axes in the order of the index columns
(x, y, z); ``padding`` is ignored for SubM convs.  Independently cross-checked against
``torch.nn.functional.conv3d`` on a densified grid in tests/test_oracle.py.
"""
from __future__ import annotations

import numpy as np
import torch


def tap_offsets(k: int) -> np.ndarray:
    """[k^3, 3] offsets; tap t = (i*k + j)*k + l  <->  delta = (i-r, j-r, l-r) on (x,y,z)."""
    r = k // 2
    ax = np.arange(-r, r + 1)
    return np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)


def kernel_map(grid_coord, batch, k: int) -> np.ndarray:
    """nbr[N, k^3] int32: index of the voxel at grid_coord[p] + delta_t in the same batch
    item, or -1.  Requires de-duplicated (batch, grid_coord)."""
    g = np.asarray(grid_coord).astype(np.int64)
    b = np.asarray(batch).astype(np.int64)
    n = g.shape[0]
    r = k // 2
    ext = g.max(0) + 1 + 2 * r

    def lin(bb, c):
        return ((bb * ext[0] + (c[:, 0] + r)) * ext[1] + (c[:, 1] + r)) * ext[2] + (c[:, 2] + r)

    base = lin(b, g)
    order = np.argsort(base, kind="stable")
    skeys = base[order]
    assert n == 0 or np.all(np.diff(skeys) != 0), "kernel_map needs de-duplicated voxels"
    offs = tap_offsets(k)
    nbr = np.full((n, k ** 3), -1, dtype=np.int32)
    for t, d in enumerate(offs):
        q = lin(b, g + d)
        pos = np.clip(np.searchsorted(skeys, q), 0, max(n - 1, 0))
        hit = skeys[pos] == q
        nbr[:, t] = np.where(hit, order[pos], -1)
    return nbr


def subm_conv(feat: torch.Tensor, nbr, weight: torch.Tensor, bias=None) -> torch.Tensor:
    """fp32 gather-GEMM over taps.  weight [Cout, k, k, k, Cin]."""
    feat = feat.float()
    cout = weight.shape[0]
    k3 = weight.shape[1] * weight.shape[2] * weight.shape[3]
    w = weight.float().reshape(cout, k3, -1)
    nbr_t = torch.as_tensor(np.asarray(nbr)).long()
    out = torch.zeros(feat.shape[0], cout, dtype=torch.float32)
    for t in range(k3):
        col = nbr_t[:, t]
        rows = (col >= 0).nonzero(as_tuple=True)[0]
        if rows.numel():
            out.index_add_(0, rows, feat[col[rows]] @ w[:, t, :].t())
    if bias is not None:
        out = out + bias.float()
    return out


def dense_conv3d_check(feat: torch.Tensor, grid_coord, weight: torch.Tensor, bias=None) -> torch.Tensor:
    """Independent check: densify one batch item, run F.conv3d (cross-correlation, zero
    padding k//2), read back at the active sites."""
    g = torch.as_tensor(np.asarray(grid_coord)).long()
    k = weight.shape[1]
    ext = (g.max(0).values + 1).tolist()
    dense = torch.zeros(1, feat.shape[1], *ext, dtype=torch.float32)
    dense[0, :, g[:, 0], g[:, 1], g[:, 2]] = feat.float().t()
    w = weight.float().permute(0, 4, 1, 2, 3).contiguous()  # [Cout, Cin, kx, ky, kz]
    out = torch.nn.functional.conv3d(dense, w, bias=None if bias is None else bias.float(), padding=k // 2)
    return out[0, :, g[:, 0], g[:, 1], g[:, 2]].t().contiguous()
