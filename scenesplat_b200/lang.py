"""Language-pretraining wrapper, losses and zero-shot head on the sm_100a kernels.

Drop-in for (reference):
  LangPretrainer                 pointcept/models/default.py:77-176
  Criteria / build_criteria      pointcept/models/losses/builder.py:13-36
  CosineSimilarity, L2Loss, AggregatedContrastiveLoss   pointcept/models/losses/misc.py:247-421
  zero-shot head                 pointcept/engines/hooks/evaluator.py:793-800, pointcept/engines/test.py:335-349
Inference uses the fused kernels; under autograd (training) the losses switch to torch operators
(scenesplat_b200/training.py).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .registry import LOSSES, MODELS, build_model
from .structure import Point


@LOSSES.register_module()
class CosineSimilarity(nn.Module):
    def __init__(self, reduction="mean", loss_weight=1.0):
        super().__init__()
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, valid_feat_mask, **kwargs):
        if pred.requires_grad:
            from . import training
            assert self.reduction == "mean"
            return training.cosine_loss(pred, target, valid_feat_mask, self.loss_weight)
        acc = ops.cos_l2_sums(pred, target, valid_feat_mask)
        loss = acc[0]
        if self.reduction == "mean":
            loss = torch.where(acc[2] > 0, acc[0] / acc[2].clamp(min=1), acc[0])
        return self.loss_weight * loss.float()


@LOSSES.register_module()
class L2Loss(nn.Module):
    def __init__(self, reduction="mean", loss_weight=1.0):
        super().__init__()
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, valid_feat_mask, **kwargs):
        if pred.requires_grad:
            from . import training
            assert self.reduction == "mean"
            return training.l2_loss(pred, target, valid_feat_mask, self.loss_weight)
        acc = ops.cos_l2_sums(pred, target, valid_feat_mask)
        loss = acc[1]
        if self.reduction == "mean":
            loss = torch.where(acc[2] > 0, acc[1] / acc[2].clamp(min=1), acc[1])
        return self.loss_weight * loss.float()


@LOSSES.register_module()
class AggregatedContrastiveLoss(nn.Module):
    def __init__(self, temperature=0.2, reduction="mean", loss_weight=1.0, schedule="all", max_classes=256):
        super().__init__()
        self.temperature, self.reduction, self.loss_weight, self.schedule = temperature, reduction, loss_weight, schedule
        self.max_classes = max_classes
        if "last_" in self.schedule:
            self.last_percent = float(self.schedule.split("_")[-1]) / 100

    @staticmethod
    def random_halves(valid, segment, n_classes):
        """Per-class random half split, one device pass (the reference loops over labels with a
        `torch.randperm` each, losses/misc.py:361-372): rank points inside their class by a random key;
        the first n//2 of each class form group a (0), the rest group b (1)."""
        n = segment.shape[0]
        lab = torch.where(valid, segment.long(), torch.full_like(segment.long(), n_classes))
        key = lab.double() + torch.rand(n, device=segment.device, dtype=torch.float64) * 0.5
        order = torch.argsort(key)
        counts = torch.bincount(lab, minlength=n_classes + 1)
        start = torch.cumsum(counts, 0) - counts
        rank = torch.empty(n, dtype=torch.long, device=segment.device)
        rank[order] = torch.arange(n, device=segment.device) - start[lab[order]]
        return (rank >= counts[lab] // 2).long(), counts[:n_classes]

    def forward(self, pred, target, valid_feat_mask, segment, epoch_progress=None, half=None, **kwargs):
        device = pred.device
        if "last_" in self.schedule and epoch_progress is not None:
            if epoch_progress <= (1 - self.last_percent):
                return torch.tensor(0.0, device=device)
        elif self.schedule == "skip":
            return torch.tensor(0.0, device=device)
        if segment is None:
            return torch.tensor(0.0, device=device)
        valid = (valid_feat_mask > 0) & (segment != -1)
        nc = self.max_classes
        if half is None:
            half, _ = self.random_halves(valid, segment, nc)
        if pred.requires_grad:  # training: differentiable and free of host decisions
            from . import training
            sums, counts = training.class_half_sums(pred.float(), valid, segment, half, nc)
            return self.loss_weight * training.contrastive_from_sums(sums, counts, nc, self.temperature, self.reduction)
        sums, counts = ops.class_half_sums(pred, valid, segment, half, nc)
        per_class = counts.view(nc, 2).sum(1)
        use = (per_class >= 100) & (counts.view(nc, 2).min(1).values > 0)  # losses/misc.py:366-376
        idx = use.nonzero(as_tuple=True)[0]
        if idx.numel() == 0:
            return torch.tensor(0.0, device=device)
        s = sums.view(nc, 2, -1)[idx]
        a = F.normalize(s[:, 0], p=2, dim=1)
        b = F.normalize(s[:, 1], p=2, dim=1)
        logits = a @ b.t() / self.temperature
        tgt = torch.arange(logits.shape[0], device=device)
        loss = (F.cross_entropy(logits, tgt) + F.cross_entropy(logits.t(), tgt)) / 2.0
        if self.reduction == "sum":
            loss = loss * logits.shape[0]
        return self.loss_weight * loss


class Criteria(object):
    """losses/builder.py:13-32"""

    def __init__(self, cfg=None):
        self.cfg = cfg if cfg is not None else []
        self.criteria = [LOSSES.build(cfg=c) for c in self.cfg]

    def __call__(self, pred, target, **kwargs):
        if len(self.criteria) == 0:
            return pred
        loss = 0
        for c in self.criteria:
            loss = loss + c(pred, target, **kwargs)
        return loss


def build_criteria(cfg):
    return Criteria(cfg)


@MODELS.register_module()
class LangPretrainer(nn.Module):
    """default.py:77-176"""

    def __init__(self, backbone=None, criteria=None):
        super().__init__()
        self.backbone = build_model(backbone)
        self.criteria = build_criteria(criteria)

    def _features(self, input_dict):
        point = Point(input_dict)
        if not (torch.is_grad_enabled() and self.training) and self._fused_normalize():
            return self.features_prepared(self.backbone.prepare(point))
        point_feat = self.backbone(point)
        if point_feat["feat"].requires_grad:  # training: torch operator under autograd
            point_feat["feat"] = F.normalize(point_feat["feat"].float(), p=2, dim=1)
        else:
            point_feat["feat"] = ops.l2_normalize(point_feat["feat"], eps=1e-12)  # F.normalize(p=2, dim=1), default.py:98
        return point_feat

    def _fused_normalize(self):
        """The backbone can write F.normalize(feat) from its last Block (PointTransformerV3.run(l2_normalize=True))."""
        return bool(getattr(self.backbone, "fused_l2_normalize", False))

    def prepare(self, input_dict):
        """Index phase (serialization, pooling levels, kernel maps; all host syncs) -> prepared Point."""
        return self.backbone.prepare(Point(input_dict))

    def features_prepared(self, point):
        """Feature phase on a prepared Point (eval): L2-normalised point features, no host sync."""
        point_feat = self.backbone.run(point, l2_normalize=True) if self._fused_normalize() else self.backbone.run(point)
        if not point_feat.pop("_l2_normalized", False):
            point_feat["feat"] = ops.l2_normalize(point_feat["feat"], eps=1e-12)
        return point_feat

    def forward(self, input_dict, chunk_size=None):
        if chunk_size is not None and chunk_size > 0 and input_dict["coord"].shape[0] > chunk_size:
            return self._chunked_forward(input_dict, chunk_size)
        point_feat = self._features(input_dict)
        if self.training:
            segment = input_dict["segment"] if "segment" in input_dict.keys() else None
            loss = self.criteria(point_feat["feat"], input_dict["lang_feat"],
                                 valid_feat_mask=input_dict["valid_feat_mask"], segment=segment,
                                 epoch_progress=input_dict["epoch_progress"])
            return dict(loss=loss)
        return dict(point_feat=point_feat)

    def _chunked_forward(self, input_dict, chunk_size):
        """default.py:115-176: contiguous index-range chunks, independent backbone passes."""
        coords = input_dict["coord"]
        N = coords.shape[0]
        outs = []
        for start in range(0, N, chunk_size):
            end = min(start + chunk_size, N)
            chunk = {k: v[start:end] for k, v in input_dict.items() if isinstance(v, torch.Tensor) and v.shape[0] == N}
            if "condition" in input_dict.keys():
                chunk["condition"] = input_dict["condition"][0]
            chunk["offset"] = torch.tensor([end - start], device=coords.device)
            pf = self._features(chunk)
            if self.training:
                outs.append(self.criteria(pf["feat"], chunk["lang_feat"], valid_feat_mask=chunk["valid_feat_mask"],
                                          segment=chunk.get("segment", None),
                                          epoch_progress=chunk.get("epoch_progress", None)))
            else:
                outs.append(pf["feat"])
        if self.training:
            return dict(loss=torch.stack(outs).mean())
        return dict(point_feat={"feat": torch.cat(outs, dim=0)})


def zero_shot_labels(feat, text_embeddings, threshold=0.1, normalize=False):
    """evaluator.py:793-800: sigmoid(feat @ T^T) -> (max prob, argmax label, -1 below threshold); fused."""
    return ops.lang_head_argmax(feat, text_embeddings, normalize=normalize, threshold=threshold)


def zero_shot_accumulate(pred_probs, feat, text_embeddings, idx_part=None):
    """test.py:335-349: pred[idx_part] += sigmoid(feat @ T^T); fused, logits never materialised."""
    return ops.lang_head_accumulate(feat, text_embeddings, pred_probs, idx_part)


class ChunkPipeline(object):
    """Two-stage software pipeline over a stream of independent chunks (the unit the path shards by: batch items,
    preprocessing chunks, test fragments; pointcept/models/default.py:134-176, pointcept/engines/test.py:315-349).

    Stage 1 (side stream): host-to-device copy of the chunk (if it comes from pinned host memory) and the index phase
    (`LangPretrainer.prepare`: serialization, pooling levels, kernel maps, pair lists; the only place with host syncs).
    Stage 2 (main stream): the feature phase.  While the device runs the heavy kernels of chunk i, the host and
    a few SMs build the indices of chunk i + 1, so the main stream never drains.  Results are identical to calling
    the model chunk by chunk (same kernels, same order of the CPU RNG draws).

        pipe = ChunkPipeline(model)
        for feat in pipe.map(chunks):   # yields the [N_i, 768] feature tensor of every chunk, in order
            ...
    """

    def __init__(self, model, device=None):
        self.model = model
        self.device = device or next(model.parameters()).device
        # high priority: the index phase is a handful of small kernels with host syncs between them; its CTAs must not
        # queue behind the feature phase's waves or stage 1 stops hiding under stage 2
        self.side = torch.cuda.Stream(device=self.device, priority=-1)
        self._keep = []  # prepared points stay referenced until the main stream has consumed them

    def _stage1(self, chunk):
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.side):
            d = {k: (v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in chunk.items()}
            for v in d.values():
                # Memory made here belongs to the side stream's allocator pool, but the feature phase reads it on
                # `main`: tell the allocator, so a block is never handed to the next chunk's copy while a kernel of
                # this chunk still has to read it (the input `feat` loses its last host reference the moment
                # Embedding.forward rebinds point.feat, long before the stem conv has run).
                if isinstance(v, torch.Tensor) and v.is_cuda:
                    v.record_stream(main)
            point = self.model.prepare(d)
            ev = torch.cuda.Event()
            ev.record(self.side)
        return point, d, ev, main

    def map(self, chunks):
        it = iter(chunks)
        try:
            nxt = self._stage1(next(it))
        except StopIteration:
            return
        while nxt is not None:
            point, inputs, ev, main = nxt
            main.wait_event(ev)
            with torch.no_grad():
                out = self.model.features_prepared(point)
            done = torch.cuda.Event()
            done.record(main)
            # the prepared Point AND the chunk's device inputs stay referenced until `done` has completed
            self._keep.append((point, inputs, done))
            try:
                nxt = self._stage1(next(it))  # overlaps the feature phase just enqueued
            except StopIteration:
                nxt = None
            while len(self._keep) > 2:  # memory made on the side stream is released only after its last use
                ent = self._keep.pop(0)
                ent[-1].synchronize()
                del ent
            yield out["feat"]

    def flush(self):
        for ent in self._keep:
            ent[-1].synchronize()
        self._keep.clear()
