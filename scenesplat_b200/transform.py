"""GridSample on the GPU -- drop-in for pointcept/datasets/transform.py:1181-1416 (same constructor
kwargs and dict keys).  The voxel hash / sort / unique / select run as sm_100a kernels; attribute
arrays are gathered on the device.  Accepts numpy arrays (returned as numpy, like the reference) or
torch tensors (returned on the GPU, which skips the host round trip when the model follows directly).

Deterministic outputs are bit-exact with the reference (`inverse`, `grid_coord`, voxel order, counts).
The representative of a voxel is member `r % count` of the voxel's members in ORIGINAL INDEX order,
with `r` drawn from numpy's global RNG exactly like transform.py:1264-1267 (the reference's own
choice additionally depends on numpy's unstable introsort, so only membership can be compared).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .registry import TRANSFORMS


@TRANSFORMS.register_module()
class GridSample(object):
    """transform.py:1211-1330.  Voxel index = floor(float64(coord) / grid_size), i.e. the reference's
    `coord / np.array(grid_size)` as NumPy >= 2 evaluates it for float32 coordinates (NumPy 1.x keeps that division in
    float32: ids at exact cell boundaries can differ; the golden fixtures were generated with NumPy 2.3)."""

    def __init__(self, grid_size=0.05, hash_type="fnv", mode="train", keys=("coord", "color", "normal", "segment"),
                 return_inverse=False, return_grid_coord=False, return_min_coord=False, return_displacement=False,
                 project_displacement=False, importance_sample_key=None, apply_to_pc=True, device="cuda"):
        assert mode in ["train", "test"]
        if importance_sample_key is not None:
            raise NotImplementedError("importance sampling is not used by the lang configs (out of scope)")
        self.grid_size, self.hash_type, self.mode, self.keys = grid_size, hash_type, mode, keys
        self.return_inverse, self.return_grid_coord = return_inverse, return_grid_coord
        self.return_min_coord, self.return_displacement = return_min_coord, return_displacement
        self.project_displacement, self.apply_to_pc = project_displacement, apply_to_pc
        self.device = device

    def _to_dev(self, v):
        if isinstance(v, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(v)).to(self.device, non_blocking=True)
        return v.to(self.device, non_blocking=True)

    def __call__(self, data_dict):
        assert "coord" in data_dict.keys()
        if "pc_coord" in data_dict and self.apply_to_pc:
            raise NotImplementedError("pc_coord co-sampling is dataset-specific preprocessing (out of scope)")
        if "sampled_index" in data_dict:
            raise NotImplementedError("sampled_index (ScanNet data-efficient) is out of scope")
        as_numpy = isinstance(data_dict["coord"], np.ndarray)
        coord = self._to_dev(data_dict["coord"]).float().contiguous()
        ix = ops.gridsample_index(coord, self.grid_size, self.hash_type)
        m = ix["m"]
        count = ix["start"][1: m + 1] - ix["start"][:m]
        back = (lambda t: t.cpu().numpy()) if as_numpy else (lambda t: t)

        def emit(idx, gc, out):
            if self.return_grid_coord:
                out["grid_coord"] = back(gc)
            if self.return_min_coord:
                mn = ix["min_coord"].double() * self.grid_size
                out["min_coord"] = back(mn.reshape(1, 3))
            if self.return_displacement:
                scaled = coord[idx].double() / self.grid_size - ix["min_coord"].double()
                disp = scaled - gc.double() - 0.5
                if self.project_displacement:
                    disp = (disp * self._to_dev(data_dict["normal"])[idx]).sum(-1, keepdim=True)
                out["displacement"] = back(disp)

        if self.mode == "train":
            cmax = int(count.max()) if m > 0 else 1
            rnd = torch.from_numpy(np.random.randint(0, cmax, m)).to(self.device)
            idx, gc, _ = ops.gridsample_select(ix, rnd, want_grid_coord=True)
            if self.return_inverse:
                data_dict["inverse"] = back(ix["inverse"])
            emit(idx, gc, data_dict)
            for key in self.keys:
                if key in data_dict.keys():
                    data_dict[key] = back(ops.gather_rows(self._to_dev(data_dict[key]), idx))
            return data_dict

        parts = []
        cmax = int(count.max()) if m > 0 else 0
        dev_cache = {k: self._to_dev(v) for k, v in data_dict.items() if k in self.keys}
        for i in range(cmax):
            idx, gc, _ = ops.gridsample_select(ix, None, frag=i, want_grid_coord=True)
            part = dict(index=back(idx))
            if self.return_inverse:
                data_dict["inverse"] = back(ix["inverse"])
            emit(idx, gc, part)
            for key in data_dict.keys():
                if key in self.keys:
                    part[key] = back(ops.gather_rows(dev_cache[key], idx))
                else:
                    part[key] = data_dict[key]
            parts.append(part)
        return parts


@TRANSFORMS.register_module()
class SphereCrop(object):
    """Drop-in for pointcept/datasets/transform.py:1419-1535, modes "random", "center" and "all" (the test-time multi-crop
    generator, ref :1439-1503).  The distance ranking runs on the GPU (csrc/crop.cu), every per-point array of the dict
    is gathered on the device.  Random numbers are drawn from numpy's global RNG exactly like the reference
    (`np.random.randint(N)`; for "all" `np.random.rand(N) * 1e-3`)."""

    # every key the reference crops (transform.py:1500-1534), plus `index` if present
    KEYS = ("coord", "origin_coord", "grid_coord", "color", "quat", "scale", "opacity", "normal", "lang_feat",
            "valid_feat_mask", "segment", "instance", "displacement", "strength")

    def __init__(self, point_max=80000, sample_rate=None, mode="random", device="cuda"):
        assert mode in ["random", "center", "all"]
        self.point_max, self.sample_rate, self.mode, self.device = point_max, sample_rate, mode, device

    # keys the reference copies into every crop of mode "all" (ref :1462-1491)
    ALL_KEYS = ("coord", "grid_coord", "normal", "color", "opacity", "quat", "lang_feat", "valid_feat_mask", "scale",
                "displacement", "strength")

    def _crop_all(self, data_dict, point_max):
        """ref :1439-1503: crops of the `point_max` nearest points around successive seeds until every point is covered;
        the next seed is the point with the smallest priority, and a crop raises the priority of its members by
        (1 - dist2 / max dist2)^2.  Returns a list of dicts (with `weight` = dist2 and `index`)."""
        n = data_dict["coord"].shape[0]
        as_numpy = isinstance(data_dict["coord"], np.ndarray)
        to_dev = lambda v: (torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v).to(self.device)
        back = (lambda t: t.cpu().numpy()) if as_numpy else (lambda t: t)
        if "index" not in data_dict.keys():
            data_dict["index"] = np.arange(n) if as_numpy else torch.arange(n, device=self.device)
        if n <= point_max:
            part = dict(data_dict)
            part["weight"] = np.zeros(n) if as_numpy else torch.zeros(n, dtype=torch.float64, device=self.device)
            return [part]
        coord = to_dev(data_dict["coord"]).float().contiguous()
        index = to_dev(data_dict["index"])
        dev = {k: to_dev(data_dict[k]) for k in self.ALL_KEYS if k in data_dict.keys()}
        prio = torch.from_numpy(np.random.rand(n) * 1e-3).to(self.device)  # float64, the reference's `coord_p`
        covered = torch.zeros(n, dtype=torch.bool, device=self.device)
        parts = []
        while not bool(covered.all()):
            init = int(torch.argmin(prio))
            center = coord[init]
            order = ops.sphere_crop_order(coord, center.cpu())
            idx_crop = order[:point_max].contiguous()
            part = {k: back(ops.gather_rows(v, idx_crop)) for k, v in dev.items()}
            d = coord[idx_crop] - center
            sq = d * d  # numpy: np.sum(np.power(coord - c, 2), 1) = ((x^2 + y^2) + z^2), every step rounded to fp32
            dist2 = (sq[:, 0] + sq[:, 1]) + sq[:, 2]
            part["weight"] = back(dist2)
            part["index"] = back(index[idx_crop])
            parts.append(part)
            # numpy: float32 `weight` -> float32 delta, added into the float64 priorities
            prio[idx_crop] += torch.square(1.0 - dist2 / dist2.max()).double()
            covered[idx_crop] = True
        return parts

    def __call__(self, data_dict):
        assert "coord" in data_dict.keys()
        n = data_dict["coord"].shape[0]
        point_max = int(self.sample_rate * n) if self.sample_rate is not None else self.point_max
        if self.mode == "all":
            return self._crop_all(data_dict, point_max)
        if n <= point_max:
            return data_dict
        as_numpy = isinstance(data_dict["coord"], np.ndarray)
        to_dev = lambda v: (torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v).to(self.device)
        coord = to_dev(data_dict["coord"]).float().contiguous()
        ci = np.random.randint(n) if self.mode == "random" else n // 2
        order = ops.sphere_crop_order(coord, coord[ci].cpu())
        idx_crop = order[:point_max].contiguous()
        for k in self.KEYS:
            if k in data_dict.keys():
                v = data_dict[k]
                g = ops.gather_rows(to_dev(v), idx_crop)
                data_dict[k] = g.cpu().numpy() if as_numpy else g
        return data_dict
