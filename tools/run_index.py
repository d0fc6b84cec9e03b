"""Stand-alone driver of the index path at the benchmark chunk's size (developer tool for timing / ncu):
Point.serialization (4 orders) + the three SerializedPooling levels (index build + feature / coordinate reduction),
each timed with CUDA events around the C-ABI call, back to back on an otherwise idle stream, and checked against torch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scenesplat_b200 import ops, synthetic, _lib as L
from oracle import gridsample as ogs

n_raw = int(os.environ.get("INDEX_NRAW", 360000))
reps = int(os.environ.get("INDEX_REPS", 20))
check = int(os.environ.get("INDEX_CHECK", 1))
dev = torch.device("cuda")
d = synthetic.chunk(n_raw, seed=0)
res = ogs.grid_sample_train(d["coord"], 0.02)
g = torch.from_numpy(res["grid_coord"]).to(dev)
coord = torch.from_numpy(d["coord"][res["idx_unique"]]).to(dev).float()
n = g.shape[0]
offset = torch.tensor([n], device=dev)
depth = ops.coord_depth(g)
orders = ("z", "z-trans", "hilbert", "hilbert-trans")
chans = (64, 128, 256)
print(f"n = {n} voxels, depth {depth}")


def once(profile):
    L.PROFILE = {} if profile else None
    _, code, order, inverse = ops.serialize(g, offset, depth, orders, want_batch=False)
    batch = torch.zeros(n, dtype=torch.int64, device=dev)
    lv = [(code, order, g.long(), batch, coord)]
    outs = []
    for c in chans:
        code_, order_, gc_, batch_, coord_ = lv[-1]
        ix = ops.pool_index(code_, order_, gc_, batch_, 1, [0, 1, 2, 3])
        src = torch.empty((code_.shape[1], c), dtype=torch.bfloat16, device=dev).normal_()
        feat, cm = ops.pool_reduce(src, coord_, order_[0].contiguous(), ix["seg_start"], "mean", out_dtype=torch.float32)
        outs.append((ix, src, feat, cm))
        lv.append((ix["code"], ix["order"], ix["grid_coord"], ix["batch"], cm))
    torch.cuda.synchronize()
    prof, L.PROFILE = L.PROFILE, None
    return (code, order, inverse), lv, outs, prof


ser, lv, outs, _ = once(False)
if check:
    code, order, inverse = ser
    ref_order = torch.argsort(code, dim=1, stable=True)
    assert torch.equal(order, ref_order), "order"
    ar = torch.arange(n, device=dev)
    for r in range(4):
        assert torch.equal(inverse[r][order[r]], ar), "inverse"
    for lvl, (ix, src, feat, cm) in enumerate(outs):
        pc, po, pgc, pb, pcoord = lv[lvl]
        c3 = pc >> 3
        uniq, cluster = torch.unique(c3[0], sorted=True, return_inverse=True)
        assert ix["m"] == uniq.numel(), "m"
        assert torch.equal(ix["cluster"], cluster), "cluster"
        assert torch.equal(ix["code"][0], uniq), "code row 0"
        ro = torch.argsort(ix["code"], dim=1, stable=True)
        assert torch.equal(ix["order"], ro), f"pooled order level {lvl}"
        for r in range(4):
            assert torch.equal(ix["inverse"][r][ix["order"][r]], torch.arange(ix["m"], device=dev)), "pooled inverse"
        want = torch.zeros((ix["m"], src.shape[1]), device=dev).index_add_(0, cluster, src.float())
        cnt = torch.bincount(cluster, minlength=ix["m"]).float()[:, None]
        assert torch.allclose(feat, want / cnt, rtol=1e-5, atol=1e-5), "segment mean"
        wc = torch.zeros((ix["m"], 3), device=dev).index_add_(0, cluster, pcoord) / cnt
        assert torch.allclose(cm, wc, rtol=1e-5, atol=1e-5), "coord mean"
    print("checks vs torch (stable argsort / unique / index_add): ok")

acc = {}
for _ in range(reps):
    _, lv, _, prof = once(True)
    for k, recs in prof.items():
        for i, (a, b, m) in enumerate(recs):
            e = acc.setdefault((k, i), dict(ms=[], bytes=(m or {}).get("bytes", 0.0)))
            e["ms"].append(a.elapsed_time(b))
tot_ms = tot_b = 0.0
for (k, i), e in acc.items():
    ms = float(np.median(e["ms"]))
    tot_ms += ms
    tot_b += e["bytes"]
    print(f"{k}[{i}]: {ms * 1e3:7.1f} us median ({min(e['ms']) * 1e3:.1f} min)  {e['bytes'] / 1e6:7.1f} MB  "
          f"{e['bytes'] / ms / 1e6:7.0f} GB/s")
print(f"total {tot_ms * 1e3:.1f} us, {tot_b / 1e6:.1f} MB algorithmic, {tot_b / tot_ms / 1e6:.0f} GB/s")
