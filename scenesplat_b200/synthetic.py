"""Seeded synthetic 3DGS chunks: a room-like surface point set
(floor / ceiling / four walls / random boxes, 1 cm Gaussian thickness) with attribute
distributions matching the reference's loaders:

  color   uint8 0..255 -> f32 / 127.5 - 1          (pointcept/datasets/transform.py:415-420)
  opacity sigmoid(N(0, 2)), shape [N, 1]           (scripts/preprocess_gs.py:50-53)
  scale   exp(N(-4, 1)) clipped to [0, 1.5]        (pointcept/datasets/scannetgs.py:114-117)
  quat    normalised N(0,1)^4 with w >= 0          (scripts/preprocess_gs.py:70-75)
  feat    = cat(color, opacity, quat, scale)       (configs/scannet/lang-pretrain-*.py:171)

numpy only; used by bench.py, the tests and smoke().
"""
from __future__ import annotations

import numpy as np


def room(n_raw: int, L: float = 6.0, H: float = 3.0, seed: int = 0, n_boxes: int = 12):
    rng = np.random.default_rng(seed)
    # surfaces: (area weight, sampler)
    surfaces = []
    surfaces.append((L * L, lambda m: np.stack([rng.uniform(0, L, m), rng.uniform(0, L, m), np.zeros(m)], 1)))
    surfaces.append((L * L, lambda m: np.stack([rng.uniform(0, L, m), rng.uniform(0, L, m), np.full(m, H)], 1)))
    surfaces.append((L * H, lambda m: np.stack([np.zeros(m), rng.uniform(0, L, m), rng.uniform(0, H, m)], 1)))
    surfaces.append((L * H, lambda m: np.stack([np.full(m, L), rng.uniform(0, L, m), rng.uniform(0, H, m)], 1)))
    surfaces.append((L * H, lambda m: np.stack([rng.uniform(0, L, m), np.zeros(m), rng.uniform(0, H, m)], 1)))
    surfaces.append((L * H, lambda m: np.stack([rng.uniform(0, L, m), np.full(m, L), rng.uniform(0, H, m)], 1)))
    for _ in range(n_boxes):
        sz = rng.uniform(0.3, 1.5, 3) * np.array([1.0, 1.0, 0.8])
        c0 = np.array([rng.uniform(0, L - sz[0]), rng.uniform(0, L - sz[1]), 0.0])
        area = 2 * (sz[0] * sz[1] + sz[1] * sz[2] + sz[0] * sz[2])

        def box(m, c0=c0, sz=sz):
            face = rng.integers(0, 6, m)
            u = rng.uniform(0, 1, (m, 3)) * sz
            ax = face // 2
            side = face % 2
            u[np.arange(m), ax] = side * sz[ax]
            return c0 + u

        surfaces.append((area, box))
    w = np.array([s[0] for s in surfaces])
    counts = rng.multinomial(n_raw, w / w.sum())
    pts = np.concatenate([s[1](int(c)) for s, c in zip(surfaces, counts) if c > 0], 0)
    pts = pts + rng.normal(0, 0.01, pts.shape)
    rng.shuffle(pts, axis=0)
    return pts.astype(np.float32)


def gaussian_attributes(n: int, seed: int = 0, lang_dim: int = 0, n_classes: int = 200):
    rng = np.random.default_rng(seed + 1)
    color = rng.integers(0, 256, (n, 3)).astype(np.float32) / 127.5 - 1.0
    opacity = (1.0 / (1.0 + np.exp(-rng.normal(0, 2, (n, 1))))).astype(np.float32)
    scale = np.clip(np.exp(rng.normal(-4, 1, (n, 3))), 0, 1.5).astype(np.float32)
    quat = rng.normal(0, 1, (n, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    quat[quat[:, 0] < 0] *= -1
    out = dict(color=color, opacity=opacity, quat=quat.astype(np.float32), scale=scale)
    if lang_dim:
        lf = rng.normal(0, 1, (n, lang_dim)).astype(np.float32)
        lf /= np.linalg.norm(lf, axis=1, keepdims=True)
        out["lang_feat"] = lf.astype(np.float16)
        out["valid_feat_mask"] = rng.uniform(0, 1, n) < 0.8
        out["segment"] = rng.integers(-1, n_classes, n).astype(np.int32)
    return out


def chunk(n_raw: int, L: float = 6.0, H: float = 3.0, seed: int = 0, lang_dim: int = 0):
    """One raw synthetic 3DGS chunk: dict(coord, color, opacity, quat, scale[, lang_feat, ...])."""
    coord = room(n_raw, L, H, seed)
    d = dict(coord=coord)
    d.update(gaussian_attributes(coord.shape[0], seed, lang_dim))
    return d


def feat_from(d):
    return np.concatenate([d["color"], d["opacity"], d["quat"], d["scale"]], 1).astype(np.float32)
