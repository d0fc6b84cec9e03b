"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the language head and the
three language-pretraining losses.  Never imported by the product package.

Parity status: PINNED against the reference's own loss modules imported from
/root/reference (tests/golden/losses.npz) and the shipped SigLIP2 text-embedding
statistics.

Follows:
  * F.normalize + criteria          pointcept/models/default.py:88-113
  * CosineSimilarity                pointcept/models/losses/misc.py:247-270
  * L2Loss                          pointcept/models/losses/misc.py:273-295
  * AggregatedContrastiveLoss       pointcept/models/losses/misc.py:298-421
  * zero-shot head                  pointcept/engines/hooks/evaluator.py:793-800,
                                    pointcept/engines/test.py:335-349,379-383
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def normalize(feat: torch.Tensor) -> torch.Tensor:
    """default.py:98  F.normalize(p=2, dim=1) (eps 1e-12)."""
    return F.normalize(feat.float(), p=2, dim=1)


def cosine_loss(pred, target, mask, loss_weight=1.0):
    """losses/misc.py:254-270, reduction='mean'.  nn.CosineSimilarity eps = 1e-8."""
    pred, target = pred.float(), target.float()
    m = mask.bool()
    cos = F.cosine_similarity(pred[m], target[m], dim=1, eps=1e-8)
    loss = (1 - cos).sum()
    n = m.sum()
    if n > 0:
        loss = loss / n
    return loss_weight * loss


def l2_loss(pred, target, mask, loss_weight=1.0):
    """losses/misc.py:280-295, reduction='mean'."""
    pred, target = pred.float(), target.float()
    m = mask.bool()
    loss = ((pred[m] - target[m]) ** 2).sum(dim=1).sum()
    n = m.sum()
    if n > 0:
        loss = loss / n
    return loss_weight * loss


def class_half_sums(pred, mask, segment, half):
    """Sum-pool of each (label, half) group for labels with >= 100 valid points
    (losses/misc.py:355-389).  ``half`` [N] in {0,1} is the explicit result of the
    reference's per-class ``torch.randperm`` split (first ``n//2`` of the permuted
    indices -> 0 = group a, rest -> 1 = group b).  Returns (labels[C], A[C,D], B[C,D])."""
    pred = pred.float()
    valid = (mask > 0) & (segment != -1)
    labs = torch.unique(segment[valid])
    A, B, used = [], [], []
    for lab in labs:
        sel = valid & (segment == lab)
        n = int(sel.sum())
        if n < 100:
            continue
        a = pred[sel & (half == 0)].sum(0)
        b = pred[sel & (half == 1)].sum(0)
        A.append(a)
        B.append(b)
        used.append(lab)
    if not A:
        return torch.empty(0, dtype=torch.long), None, None
    return torch.stack(used), torch.stack(A), torch.stack(B)


def contrastive_from_sums(A, B, temperature=0.2, loss_weight=1.0):
    """losses/misc.py:395-421, reduction='mean'."""
    if A is None:
        return torch.tensor(0.0)
    a = F.normalize(A.float(), p=2, dim=1)
    b = F.normalize(B.float(), p=2, dim=1)
    logits = a @ b.t() / temperature
    tgt = torch.arange(logits.shape[0])
    loss = (F.cross_entropy(logits, tgt) + F.cross_entropy(logits.t(), tgt)) / 2.0
    return loss_weight * loss


def zero_shot_head(feat, text_emb, threshold=0.1):
    """evaluator.py:793-800: logits = feat @ T^T; probs = sigmoid; max/argmax; label -1
    when max prob < threshold.  Returns (probs[N,K], max_prob[N], label[N])."""
    logits = feat.float() @ text_emb.float().t()
    probs = torch.sigmoid(logits)
    mx, arg = probs.max(dim=1)
    label = torch.where(mx < threshold, torch.full_like(arg, -1), arg)
    return probs, mx, label
