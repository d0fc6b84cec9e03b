"""Zero-shot evaluation tail on the GPU: neighbour voting and confusion matrix.

Mirrors pointcept/utils/misc.py:54-95 (`neighbor_voting`: cKDTree k-NN + numba majority vote, both on the host in
the reference) and the per-point loop of pointcept/engines/hooks/evaluator.py:830-834.  Same argument names and
meaning; tensors live on the GPU and the kernels are reached through the C-ABI (csrc/voting.cu).
"""
from __future__ import annotations

import math

import torch

from . import _lib as L

__all__ = ["neighbor_voting", "clustering_voting", "confusion_update"]


def neighbor_voting(coords, pred, vote_k, ignore_label, num_classes, valid_mask=None, query_coords=None, cell=None):
    """coords [N,3] float, pred [N] int, valid_mask [N] bool or None, query_coords [M,3] or None -> int32 [M or N].
    Labels of the vote_k nearest USED points (coords[valid_mask]) are majority-voted per query; ties go to the
    smallest label, a query whose neighbours carry no valid label gets ignore_label."""
    if not coords.is_cuda:
        raise L.CudaKernelError("scenesplat_b200.neighbor_voting needs CUDA tensors (there is no CPU path)")
    coords = coords.float().contiguous()
    query = coords if query_coords is None else query_coords.to(coords.device).float().contiguous()
    pred = pred.to(coords.device)
    if valid_mask is not None:
        vm = valid_mask.to(coords.device).bool()
        used, used_labels = coords[vm].contiguous(), pred[vm]
    else:
        used, used_labels = coords, pred
    m = used.shape[0]
    if m == 0:
        return pred  # reference: nothing to vote with
    used_labels = used_labels.to(torch.int32).contiguous()
    k = min(int(vote_k), m)
    lo, hi = used.amin(0), used.amax(0)
    ext = (hi - lo).double().cpu()
    origin = lo.cpu().contiguous()
    if cell is None:
        # ~ 2 (k / density)^(1/3) for a volume, generous for surfaces; only speed depends on it
        vol = float(torch.clamp(ext, min=1e-6).prod())
        cell = max(2.0 * (vol * k / max(m, 1)) ** (1.0 / 3.0) / 2.0, 1e-6)
    dims = [int(math.floor(float(e) / cell)) + 1 for e in ext]
    while dims[0] * dims[1] * dims[2] > (1 << 26):  # keep the cell table small
        cell *= 1.26
        dims = [int(math.floor(float(e) / cell)) + 1 for e in ext]
    nx, ny, nz = dims
    ncell = nx * ny * nz
    dev = coords.device
    count = torch.zeros(ncell, dtype=torch.int32, device=dev)
    cell_id = torch.empty(m, dtype=torch.int64, device=dev)
    o = L.float_array(origin.tolist())
    L.call("ss_vote_bin_count", L.ptr(used), m, o, float(cell), nx, ny, nz, L.ptr(count), L.ptr(cell_id), L.stream())
    start = torch.zeros(ncell + 1, dtype=torch.int64, device=dev)
    torch.cumsum(count, 0, out=start[1:])
    cursor = torch.zeros(ncell, dtype=torch.int32, device=dev)
    binned = torch.empty((m, 4), dtype=torch.float32, device=dev)
    L.call("ss_vote_bin_fill", L.ptr(used), L.ptr(used_labels), L.ptr(cell_id), m, L.ptr(start), L.ptr(cursor),
           L.ptr(binned), L.stream())
    out = torch.empty(query.shape[0], dtype=torch.int32, device=dev)
    L.call("ss_knn_vote", L.ptr(binned), L.ptr(start), o, float(cell), nx, ny, nz, L.ptr(query), query.shape[0], k,
           int(ignore_label), int(num_classes), L.ptr(out), L.stream())
    return out


def confusion_update(gt, pred, num_classes, ignore_index, confusion, fn_ignore):
    """confusion[gt, pred] += 1, or fn_ignore[gt] += 1 where pred == ignore_index; in place on int64 CUDA tensors.
    `gt` must already be filtered to valid labels (evaluator.py:786-787)."""
    if not confusion.is_cuda:
        raise L.CudaKernelError("scenesplat_b200.confusion_update needs CUDA tensors (there is no CPU path)")
    gt = gt.to(confusion.device, torch.int64).contiguous()
    pred = pred.to(confusion.device, torch.int64).contiguous()
    assert confusion.dtype == torch.int64 and fn_ignore.dtype == torch.int64 and confusion.is_contiguous()
    L.call("ss_confusion_update", L.ptr(gt), L.ptr(pred), gt.numel(), int(num_classes), int(ignore_index),
           L.ptr(confusion), L.ptr(fn_ignore), L.stream())
    return confusion, fn_ignore


def clustering_voting(pred, instance_labels, ignore_index):
    """pointcept/utils/misc.py:98-125 without the host loop over instances: every point of an instance (instance id !=
    ignore_index) takes the instance's most frequent predicted label; ties go to the smallest label value (the
    reference takes np.argmax over np.unique's ascending classes; the ignore label is a label value like any other).
    One 2-D histogram (instances x label values) built with a device bincount, one argmax.  pred / instance_labels:
    integer tensors [N] on the same device -> tensor like pred."""
    if pred.shape != instance_labels.shape:
        print("clustering_voting: prediction and instance arrays must have the same shape")
        return pred
    if pred.numel() == 0:
        return pred.clone()
    inst_vals, inst_idx = torch.unique(instance_labels, sorted=True, return_inverse=True)
    cls_vals, cls_idx = torch.unique(pred, sorted=True, return_inverse=True)
    ni, nc = inst_vals.numel(), cls_vals.numel()
    hist = torch.bincount(inst_idx * nc + cls_idx, minlength=ni * nc).view(ni, nc)
    major = cls_vals[torch.argmax(hist, dim=1)]          # first maximum = smallest label value
    out = major[inst_idx]
    return torch.where(instance_labels != ignore_index, out, pred)
