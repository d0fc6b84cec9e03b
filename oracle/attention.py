"""TEST INFRASTRUCTURE ONLY -- CPU restatement of SerializedAttention's patch
table and patch-wise softmax attention
(pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py:114-222).
Never imported by the product package.

Parity status: the patch table (``pad`` / ``unpad`` / ``cu_seqlens``) is PINNED
against the reference's own ``get_padding_and_inverse`` (tests/golden/patch_table.npz).
The attention arithmetic in the reference's flash branch lives in
``flash_attn.flash_attn_varlen_qkvpacked_func`` (third party, flash-attn 2,
``env.yaml:52``), absent from /root/reference -> "parity unpinned" by the reference;
this oracle restates it as plain fp32 softmax(q k^T * scale) v per
``cu_seqlens`` sequence, which is exactly the reference's own non-flash branch
(ptv3:190-206) applied to the flash-path patch table (ptv3:114-170).
"""
from __future__ import annotations

import numpy as np
import torch


def patch_table(offset, K: int):
    """ptv3:114-170.  Returns (pad[T], unpad[N], cu_seqlens[S+1] int32)."""
    offset = np.asarray(offset).astype(np.int64)
    bincount = np.diff(np.concatenate([[0], offset]))
    bincount_pad = ((bincount + K - 1) // K) * K
    mask_pad = bincount > K
    bincount_pad = np.where(mask_pad, bincount_pad, bincount)
    _offset = np.concatenate([[0], offset])
    _offset_pad = np.concatenate([[0], np.cumsum(bincount_pad)])
    pad = np.arange(_offset_pad[-1], dtype=np.int64)
    unpad = np.arange(_offset[-1], dtype=np.int64)
    cu = []
    for i in range(len(offset)):
        unpad[_offset[i]:_offset[i + 1]] += _offset_pad[i] - _offset[i]
        if bincount[i] != bincount_pad[i]:
            r = bincount[i] % K
            pad[_offset_pad[i + 1] - K + r:_offset_pad[i + 1]] = \
                pad[_offset_pad[i + 1] - 2 * K + r:_offset_pad[i + 1] - K]
        pad[_offset_pad[i]:_offset_pad[i + 1]] -= _offset_pad[i] - _offset[i]
        cu.append(np.arange(_offset_pad[i], _offset_pad[i + 1], K, dtype=np.int32))
    cu = np.concatenate(cu + [np.array([_offset_pad[-1]], dtype=np.int32)]).astype(np.int32)
    return pad, unpad, cu


def varlen_attention(qkv: torch.Tensor, cu_seqlens, H: int, scale: float) -> torch.Tensor:
    """fp32 softmax attention per sequence; qkv [T, 3*C] laid out (3, H, d) per row
    (ptv3:209 ``reshape(-1, 3, H, C // H)``).  Returns [T, C]."""
    T, C3 = qkv.shape
    C = C3 // 3
    d = C // H
    q, k, v = qkv.float().reshape(T, 3, H, d).unbind(1)
    out = torch.empty(T, H, d, dtype=torch.float32)
    cu = [int(c) for c in cu_seqlens]
    for s, e in zip(cu[:-1], cu[1:]):
        qs = q[s:e].transpose(0, 1)  # [H, L, d]
        ks = k[s:e].transpose(0, 1)
        vs = v[s:e].transpose(0, 1)
        att = torch.softmax((qs * scale) @ ks.transpose(-2, -1), dim=-1)
        out[s:e] = (att @ vs).transpose(0, 1)
    return out.reshape(T, C)


def serialized_attention_core(qkv: torch.Tensor, order_row, inverse_row, offset, K: int, H: int,
                              scale: float) -> torch.Tensor:
    """ptv3:181-216 without the qkv / proj Linear layers:
    ``qkv[order[pad]]`` -> per-patch attention -> ``[unpad[inverse]]``."""
    pad, unpad, cu = patch_table(offset, K)
    order = np.asarray(order_row)[pad]
    inverse = unpad[np.asarray(inverse_row)]
    feat = varlen_attention(qkv[torch.from_numpy(order)], cu, H, scale)
    return feat[torch.from_numpy(inverse)]
