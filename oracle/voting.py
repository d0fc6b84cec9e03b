"""CPU restatement (TEST INFRASTRUCTURE ONLY) of the reference's zero-shot evaluation tail.

neighbor_voting: pointcept/utils/misc.py:54-95 (scipy cKDTree k-NN, the same library call as the reference) with the
numba majority vote of :17-51 restated in numpy: counts over valid labels, first (lowest) class with the largest
count wins, no valid label -> ignore_label.  confusion_update: pointcept/engines/hooks/evaluator.py:830-834.
The reference functions themselves cannot run here (numba is absent from the image), so this file is pinned only
against their text: "parity unpinned" for the vote, the k-NN is the identical scipy call.
"""
import numpy as np
from scipy.spatial import cKDTree


def majority_vote(neighbor_labels, ignore_label, num_classes):
    n, k = neighbor_labels.shape
    out = np.full(n, ignore_label, dtype=np.int32)
    valid = (neighbor_labels != ignore_label) & (neighbor_labels >= 0) & (neighbor_labels < num_classes)
    for i in range(n):
        lab = neighbor_labels[i][valid[i]]
        if lab.size:
            counts = np.bincount(lab, minlength=num_classes)
            out[i] = int(np.argmax(counts))  # first maximum = the strict `>` scan of the reference
    return out


def neighbor_voting(coords, pred, vote_k, ignore_label, num_classes, valid_mask=None, query_coords=None):
    query_pts = coords if query_coords is None else query_coords
    if valid_mask is not None:
        used_coords, used_labels = coords[valid_mask], pred[valid_mask]
    else:
        used_coords, used_labels = coords, pred
    if len(used_coords) == 0:
        return pred
    _, nn = cKDTree(used_coords).query(query_pts, k=vote_k)
    if vote_k == 1:
        nn = nn[:, None]
    return majority_vote(used_labels[nn], ignore_label, num_classes)


def confusion_update(gt, pred, num_classes, ignore_index, confusion, fn_ignore):
    for g, p in zip(gt, pred):
        if p == ignore_index:
            fn_ignore[g] += 1
        else:
            confusion[g, p] += 1
    return confusion, fn_ignore


def clustering_voting(pred, instance_labels, ignore_index):
    """pointcept/utils/misc.py:98-125, restated line by line (numpy only)."""
    updated = pred.copy()
    for inst in np.unique(instance_labels):
        if inst == ignore_index:
            continue
        mask = instance_labels == inst
        classes, counts = np.unique(pred[mask], return_counts=True)
        updated[mask] = classes[np.argmax(counts)]
    return updated
