"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference's
``PointTransformerV3.forward`` in eval mode
(pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py:699-714 and the
modules it drives: Embedding :485-515, Block :251-338, SerializedAttention :51-222,
MLP :225-248, SerializedPooling :341-444, SerializedUnpooling :447-482), operating on a
plain ``state_dict`` with the reference's key names.  Never imported by the product
package.

Parity status: PINNED against the reference's own module (imported unmodified with
the shims in oracle/ref_shim.py, ``enable_flash=False`` is avoided by comparing at
sizes where every batch item has >= patch_size tokens or by using the oracle's own patch
table; see tests/golden/make_golden.py -> ptv3_small.npz).  The two third-party
operators underneath (spconv, flash_attn) are restated in oracle/subm_conv.py and
oracle/attention.py and carry their "parity unpinned" status.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import attention as oattn
from . import pooling as opool
from . import serialization as oser
from . import subm_conv as oconv


BN_TRAINING = False  # set by ptv3_forward_autograd(bn_training=True): batch statistics (train mode), momentum 0.01


def _bn_eval(x, sd, p, eps=1e-3):
    if BN_TRAINING:  # nn.BatchNorm1d(eps=1e-3, momentum=0.01).train(): running stats updated on private copies
        return F.batch_norm(x, sd[p + "running_mean"].clone(), sd[p + "running_var"].clone(), sd[p + "weight"],
                            sd[p + "bias"], training=True, momentum=0.01, eps=eps)
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        training=False, eps=eps)


def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[1],), sd[p + "weight"], sd[p + "bias"], 1e-5)


def _lin(x, sd, p):
    return F.linear(x, sd[p + "weight"], sd[p + "bias"])


class PointState:
    def __init__(self, feat, coord, grid_coord, batch, offset, code, order, inverse, depth):
        self.feat, self.coord, self.grid_coord, self.batch, self.offset = feat, coord, grid_coord, batch, offset
        self.code, self.order, self.inverse, self.depth = code, order, inverse, depth
        self.kmaps = {}
        self.parent = None
        self.cluster = None
        # features held by ``sparse_conv_feat`` when they differ from ``feat`` (see unpooling_forward)
        self.conv_feat = None


def block_forward(pt: PointState, sd, p, C, H, K, order_index, stage, taps=None):
    """ptv3:318-338 (eval: DropPath = identity)."""
    if 3 not in pt.kmaps:
        pt.kmaps[3] = oconv.kernel_map(pt.grid_coord, pt.batch, 3)
    x = pt.feat
    conv_in = x if pt.conv_feat is None else pt.conv_feat
    pt.conv_feat = None
    y = oconv.subm_conv(conv_in, pt.kmaps[3], sd[p + "cpe.0.weight"], sd[p + "cpe.0.bias"])
    y = _ln(_lin(y, sd, p + "cpe.1."), sd, p + "cpe.2.")
    x = x + y
    h = _ln(x, sd, p + "norm1.0.")
    qkv = _lin(h, sd, p + "attn.qkv.")
    a = oattn.serialized_attention_core(qkv, pt.order[order_index], pt.inverse[order_index], pt.offset, K, H,
                                        (C // H) ** -0.5)
    x = x + _lin(a, sd, p + "attn.proj.")
    h = _ln(x, sd, p + "norm2.0.")
    m = _lin(F.gelu(_lin(h, sd, p + "mlp.0.fc1.")), sd, p + "mlp.0.fc2.")
    pt.feat = x + m
    if taps is not None:
        taps[p.rstrip(".")] = pt.feat.clone()
    return pt


def pooling_forward(pt: PointState, sd, p, perm):
    """ptv3:371-444 (stride 2 -> pooling_depth 1)."""
    ix = opool.pool_index(pt.code, 1, perm)
    proj = _lin(pt.feat, sd, p + "proj.")
    if proj.requires_grad:  # differentiable twin of segment_csr(mean) for the gradient oracle
        cnt = np.diff(ix["idx_ptr"])
        seg = torch.from_numpy(np.repeat(np.arange(len(cnt)), cnt))
        feat = torch.zeros(len(cnt), proj.shape[1]).index_add_(0, seg, proj[torch.from_numpy(ix["indices"])])
        feat = feat / torch.from_numpy(cnt.astype(np.float32))[:, None]
    else:
        feat = torch.from_numpy(opool.segment_csr(proj.numpy(), ix["indices"], ix["idx_ptr"], "mean"))
    coord = opool.segment_csr(pt.coord, ix["indices"], ix["idx_ptr"], "mean")
    gc, b = opool.pooled_attrs(pt.grid_coord, pt.batch, ix["head"], 1)
    offset = np.cumsum(np.bincount(b, minlength=len(pt.offset))).astype(np.int64)
    child = PointState(feat, coord, gc, b, offset, ix["code"], ix["order"], ix["inverse"], pt.depth - 1)
    child.parent, child.cluster = pt, ix["cluster"]
    child.feat = F.gelu(_bn_eval(child.feat, sd, p + "norm.0."))
    return child


def unpooling_forward(pt: PointState, sd, p):
    """ptv3:471-482.

    REFERENCE QUIRK (defines behaviour): ``parent.feat = parent.feat + point.feat[inverse]``
    (ptv3:478) rebinds ``parent.feat`` only; ``parent.sparse_conv_feat`` still holds the
    output of ``proj_skip`` (last refreshed by PointSequential's torch-module branch,
    pointcept/models/modules.py:79-84).  The next module to run is the first decoder Block's
    ``cpe`` conv, which reads ``sparse_conv_feat.features`` (modules.py:68-72) -> the xCPE conv
    of the FIRST block after every unpooling sees the skip projection WITHOUT the unpooled
    coarse features, while the residual shortcut uses the sum."""
    parent = pt.parent
    a = F.gelu(_bn_eval(_lin(pt.feat, sd, p + "proj.0."), sd, p + "proj.1."))
    s = F.gelu(_bn_eval(_lin(parent.feat, sd, p + "proj_skip.0."), sd, p + "proj_skip.1."))
    parent.feat = s + a[torch.from_numpy(pt.cluster)]
    parent.conv_feat = s
    return parent


def ptv3_forward_autograd(sd, cfg, coord, grid_coord, feat, offset, perms=None, prefix="", bn_training=True):
    """The same forward with autograd enabled (parameters in `sd` may require grad) and, by default, train-mode
    BatchNorm: the gradient oracle of tests/test_gpu_train.py (DropPath must be off: it is random)."""
    global BN_TRAINING
    old, BN_TRAINING = BN_TRAINING, bn_training
    try:
        with torch.enable_grad():
            return _ptv3_forward(sd, cfg, coord, grid_coord, feat, offset, perms, prefix, None)
    finally:
        BN_TRAINING = old


@torch.no_grad()
def ptv3_forward(sd, cfg, coord, grid_coord, feat, offset, perms=None, prefix="", taps=None):
    return _ptv3_forward(sd, cfg, coord, grid_coord, feat, offset, perms, prefix, taps)


def _ptv3_forward(sd, cfg, coord, grid_coord, feat, offset, perms=None, prefix="", taps=None):
    """sd: fp32 state_dict with the reference's keys (optionally under ``prefix``);
    cfg: dict with the PT-v3m1 kwargs; perms: list of row permutations, one per
    ``randperm`` call in forward order (serialization, then each SerializedPooling), or
    None when shuffle_orders=False.  Returns the final [N, dec_channels[0]] fp32 feature."""
    sd = {k[len(prefix):]: v.float() for k, v in sd.items() if k.startswith(prefix) and v.is_floating_point()}
    grid_coord = np.asarray(grid_coord).astype(np.int64)
    offset = np.asarray(offset).astype(np.int64)
    batch = oser.offset2batch(offset)
    perms = list(perms) if perms is not None else None
    nxt = (lambda: perms.pop(0)) if perms is not None else (lambda: None)
    code, order, inverse, depth = oser.serialization(grid_coord, batch, len(offset), cfg["order"], perm=nxt())
    pt = PointState(torch.as_tensor(feat).float(), np.asarray(coord, dtype=np.float32), grid_coord, batch, offset,
                    code, order, inverse, depth)
    # Embedding (ptv3:485-515): SubMConv3d(5^3, no bias) -> BN -> GELU
    pt.kmaps[5] = oconv.kernel_map(pt.grid_coord, pt.batch, 5)
    x = oconv.subm_conv(pt.feat, pt.kmaps[5], sd["embedding.stem.conv.weight"], None)
    pt.feat = F.gelu(_bn_eval(x, sd, "embedding.stem.norm."))
    if taps is not None:
        taps["embedding"] = pt.feat.clone()
    n_order = len(cfg["order"])
    ns = len(cfg["enc_depths"])
    for s in range(ns):
        if s > 0:
            pt = pooling_forward(pt, sd, f"enc.enc{s}.down.", nxt())
            if taps is not None:
                taps[f"enc.enc{s}.down"] = pt.feat.clone()
        for i in range(cfg["enc_depths"][s]):
            pt = block_forward(pt, sd, f"enc.enc{s}.block{i}.", cfg["enc_channels"][s], cfg["enc_num_head"][s],
                               cfg["enc_patch_size"][s], i % n_order, s, taps)
    for s in reversed(range(ns - 1)):
        pt = unpooling_forward(pt, sd, f"dec.dec{s}.up.")
        if taps is not None:
            taps[f"dec.dec{s}.up"] = pt.feat.clone()
        for i in range(cfg["dec_depths"][s]):
            pt = block_forward(pt, sd, f"dec.dec{s}.block{i}.", cfg["dec_channels"][s], cfg["dec_num_head"][s],
                               cfg["dec_patch_size"][s], i % n_order, s, taps)
    return pt.feat
