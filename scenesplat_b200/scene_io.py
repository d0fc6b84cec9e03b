"""Packed scene files: the feed side of the hot path.

The reference stores a scene as one `.npy` per attribute and its dataset reads them one by one with `np.load`,
converts dtypes on the host and hands pageable numpy arrays to the transforms
(pointcept/datasets/scannetgs.py:59-167).  Here a scene is ONE file:

    [ 4 KiB-aligned header: magic "SSPK0001", uint64 header bytes, JSON table {name, dtype, shape, offset, nbytes} ]
    [ sections, each starting on a 4 KiB boundary, already in the dtypes `get_data` produces ]

so a load is one `readinto` of the payload into ONE pinned host buffer and one host-to-device copy of that buffer;
the attributes are views of it on either side (no per-attribute allocation, no dtype conversion at load time,
`lang_feat` stays fp16 until a kernel reads it).  `pack_scene` applies the reference's `get_data` normalisation
once, at packing time: coord / color / normal / quat / sh float32, opacity float32 [N, 1], scale float32 clipped
to [0, 1.5], lang_feat float16, valid_feat_mask bool, segment / instance int32 (segment20 or segment200 -> segment,
missing -> -1).
"""
from __future__ import annotations

import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

MAGIC = b"SSPK0001"
ALIGN = 4096
_PIECE = 32 << 20
VALID_ASSETS = ("coord", "color", "normal", "segment20", "segment200", "instance", "quat", "scale", "opacity", "sh",
                "lang_feat", "valid_feat_mask", "pc_instance", "pc_coord", "pc_segment20", "pc_segment200")
_TORCH = {"float32": torch.float32, "float16": torch.float16, "int32": torch.int32, "int64": torch.int64,
          "bool": torch.bool, "uint8": torch.uint8, "float64": torch.float64, "int16": torch.int16}


def _round_up(x, a=ALIGN):
    return (x + a - 1) // a * a


def normalize_assets(raw: dict) -> dict:
    """The dtype / shape normalisation of ScanNetGSDataset.get_data (scannetgs.py:93-150) on a dict of numpy arrays."""
    d = dict(raw)
    for k in ("coord", "pc_coord", "color", "normal", "quat", "sh"):
        if k in d:
            d[k] = np.asarray(d[k]).astype(np.float32)
    if "opacity" in d:
        d["opacity"] = np.asarray(d["opacity"]).astype(np.float32).reshape(-1, 1)
    if "scale" in d:
        d["scale"] = np.asarray(d["scale"]).astype(np.float32).clip(0, 1.5)
    if "lang_feat" in d:
        d["lang_feat"] = np.asarray(d["lang_feat"]).astype(np.float16)
    if "valid_feat_mask" in d:
        d["valid_feat_mask"] = np.asarray(d["valid_feat_mask"]).astype(bool)
    n = d["coord"].shape[0] if "coord" in d else None
    for dst, srcs in (("segment", ("segment20", "segment200")), ("pc_segment", ("pc_segment20", "pc_segment200"))):
        for s in srcs:
            if s in d:
                v = np.asarray(d[s]).reshape([-1]).astype(np.int32)
                for s2 in srcs:
                    d.pop(s2, None)
                d[dst] = v
                break
        else:
            if dst == "segment" and n is not None:
                d["segment"] = np.ones(n, dtype=np.int32) * -1
    if "instance" in d:
        d["instance"] = np.asarray(d["instance"]).reshape([-1]).astype(np.int32)
    elif n is not None:
        d["instance"] = np.ones(n, dtype=np.int32) * -1
    return d


def read_scene_folder(scene_dir: str) -> dict:
    """`.npy`-per-attribute folder -> normalised dict (what the reference's get_data returns, minus `name`)."""
    raw = {}
    for asset in sorted(os.listdir(scene_dir)):
        if asset.endswith(".npy") and asset[:-4] in VALID_ASSETS:
            raw[asset[:-4]] = np.load(os.path.join(scene_dir, asset))
    return normalize_assets(raw)


def pack_arrays(arrays: dict, out_path: str) -> int:
    """Write a dict of numpy arrays as one packed scene file.  Returns the file size."""
    table, off = [], 0
    items = [(k, np.ascontiguousarray(v)) for k, v in arrays.items() if isinstance(v, np.ndarray)]
    for k, v in items:
        if v.dtype.name not in _TORCH:
            raise ValueError(f"{k}: dtype {v.dtype} is not supported by the packed format")
        table.append(dict(name=k, dtype=v.dtype.name, shape=list(v.shape), offset=off, nbytes=int(v.nbytes)))
        off = _round_up(off + v.nbytes)
    meta = json.dumps(dict(version=1, payload_bytes=off, arrays=table)).encode()
    header = _round_up(len(MAGIC) + 8 + len(meta))
    tmp = out_path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        f.write(np.uint64(len(meta)).tobytes())
        f.write(meta)
        f.write(b"\0" * (header - len(MAGIC) - 8 - len(meta)))
        for ent, (_, v) in zip(table, items):
            f.seek(header + ent["offset"])
            f.write(memoryview(v).cast("B"))
        f.truncate(header + off)
    os.replace(tmp, out_path)
    return header + off


def pack_scene(scene_dir: str, out_path: str) -> int:
    """Reference scene folder -> packed file (normalisation applied once, here)."""
    return pack_arrays(read_scene_folder(scene_dir), out_path)


def read_header(path: str):
    with open(path, "rb") as f:
        head = f.read(len(MAGIC) + 8)
        if head[:len(MAGIC)] != MAGIC:
            raise ValueError(f"{path}: not a packed scene file")
        mlen = int(np.frombuffer(head[len(MAGIC):], dtype=np.uint64)[0])
        meta = json.loads(f.read(mlen).decode())
    return _round_up(len(MAGIC) + 8 + mlen), meta


class PackedScene:
    """One loaded scene: `host` is the single (pinned when CUDA is present) byte buffer holding every attribute,
    `arrays` are torch views of it; `to_device` moves the whole scene with one copy."""

    def __init__(self, host: torch.Tensor, table: list):
        self.host, self.table = host, table
        self.arrays = self._views(host)

    def _views(self, buf):
        out = {}
        for ent in self.table:
            t = buf[ent["offset"]:ent["offset"] + ent["nbytes"]]
            dt = _TORCH[ent["dtype"]]
            out[ent["name"]] = (t.view(dt) if ent["nbytes"] else torch.empty(0, dtype=dt, device=buf.device)).reshape(ent["shape"])
        return out

    def numpy(self) -> dict:
        return {k: v.numpy() for k, v in self.arrays.items()}

    def to_device(self, device, non_blocking=True) -> dict:
        """-> dict of device tensors: ONE host-to-device copy of the payload, attributes are views of it."""
        dev = self.host.to(device, non_blocking=non_blocking)
        return self._views(dev)


def load_scene(path: str, keys=None, pinned=None, threads: int = 8) -> PackedScene:
    """Read a packed scene into one host buffer (pinned by default when CUDA is available).  `keys` restricts the
    read to the byte range spanned by the named attributes (e.g. skip `lang_feat` for inference)."""
    header, meta = read_header(path)
    table = meta["arrays"] if keys is None else [e for e in meta["arrays"] if e["name"] in set(keys)]
    if keys is not None and len(table) != len(set(keys)):
        missing = set(keys) - {e["name"] for e in table}
        raise KeyError(f"{path}: no attribute(s) {sorted(missing)}")
    if not table:
        return PackedScene(torch.empty(0, dtype=torch.uint8), [])
    pinned = torch.cuda.is_available() if pinned is None else pinned
    if keys is None:   # everything: one contiguous read of the payload
        runs = [(0, meta["payload_bytes"], 0)]                      # (file offset, bytes, buffer offset)
        total, out_table = meta["payload_bytes"], table
    else:              # the named sections only, compacted (256-byte aligned) in the host buffer
        runs, out_table, total = [], [], 0
        for e in table:
            runs.append((e["offset"], e["nbytes"], total))
            out_table.append(dict(e, offset=total))
            total = _round_up(total + e["nbytes"], 256)
    host = torch.empty(total, dtype=torch.uint8, pin_memory=bool(pinned))
    buf = memoryview(host.numpy())
    # split the runs into <= 32 MiB pieces and read them with a few threads (os.preadv releases the GIL; one thread
    # copies from the page cache at ~7 GB/s, which would otherwise bound the loader)
    pieces = []
    for src, nbytes, dst in runs:
        for o in range(0, nbytes, _PIECE):
            pieces.append((src + o, min(_PIECE, nbytes - o), dst + o))
    fd = os.open(path, os.O_RDONLY)
    try:
        def read_piece(piece):
            src, nbytes, dst = piece
            got = 0
            while got < nbytes:
                r = os.preadv(fd, [buf[dst + got:dst + nbytes]], header + src + got)
                if not r:
                    raise IOError(f"{path}: truncated payload ({got} of {nbytes} bytes at offset {src})")
                got += r
        if len(pieces) > 1 and threads > 1:
            with ThreadPoolExecutor(max_workers=min(threads, len(pieces))) as ex:
                list(ex.map(read_piece, pieces))
        else:
            for piece in pieces:
                read_piece(piece)
    finally:
        os.close(fd)
    return PackedScene(host, out_table)
