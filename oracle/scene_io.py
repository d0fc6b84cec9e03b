"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's per-scene loading
(pointcept/datasets/scannetgs.py:59-150, ScanNetGSDataset.get_data): one np.load per `.npy` asset of the scene folder
followed by the dtype / shape normalisation, written independently of scenesplat_b200/scene_io.py."""
import os

import numpy as np

VALID_ASSETS = ["coord", "color", "normal", "segment20", "instance", "quat", "scale", "opacity", "lang_feat",
                "valid_feat_mask", "pc_instance"]          # scannetgs.py:20-32
EVAL_PC_ASSETS = ["pc_coord", "pc_segment20"]               # scannetgs.py:34


def get_data(data_path, is_train=True):
    d = {}
    for asset in os.listdir(data_path):                      # :69-81
        if not asset.endswith(".npy"):
            continue
        if asset[:-4] not in VALID_ASSETS and (is_train or asset[:-4] not in EVAL_PC_ASSETS):
            continue
        d[asset[:-4]] = np.load(os.path.join(data_path, asset))
    if "coord" in d:                                         # :93-96
        d["coord"] = d["coord"].astype(np.float32)
    if "pc_coord" in d:
        d["pc_coord"] = d["pc_coord"].astype(np.float32)
    if "pc_segment20" in d:                                  # :100-101
        d["pc_segment20"] = d["pc_segment20"].astype(np.int32)
    for k in ("color", "normal", "quat"):                    # :103-111
        if k in d:
            d[k] = d[k].astype(np.float32)
    if "opacity" in d:
        d["opacity"] = d["opacity"].astype(np.float32).reshape(-1, 1)
    if "scale" in d:                                         # :114-117
        d["scale"] = d["scale"].astype(np.float32).clip(0, 1.5)
    if "lang_feat" in d:                                     # :119-122
        d["lang_feat"] = d["lang_feat"].astype(np.float16)
    if "valid_feat_mask" in d:
        d["valid_feat_mask"] = d["valid_feat_mask"].astype(bool)
    if "segment20" in d:                                     # :124-135
        d["segment"] = d.pop("segment20").reshape([-1]).astype(np.int32)
    else:
        d["segment"] = np.ones(d["coord"].shape[0], dtype=np.int32) * -1
    if "pc_segment20" in d:                                  # :137-140
        d["pc_segment"] = d.pop("pc_segment20").reshape([-1]).astype(np.int32)
    if "instance" in d:                                      # :146-153
        d["instance"] = d.pop("instance").reshape([-1]).astype(np.int32)
    else:
        d["instance"] = np.ones(d["coord"].shape[0], dtype=np.int32) * -1
    return d
