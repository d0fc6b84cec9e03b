"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shim.py).  Runs in the build container
only (the reference tree does not exist on the GPU box); the resulting .npz files are
committed.  Usage:  python tests/golden/make_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_shim import REFERENCE_ROOT, load_reference  # noqa: E402
from scenesplat_b200 import synthetic  # noqa: E402

ORDERS = ("z", "z-trans", "hilbert", "hilbert-trans")

SMALL_CFG = dict(
    in_channels=11, order=ORDERS, stride=(2, 2, 2),
    enc_depths=(1, 1, 1, 2), enc_channels=(16, 32, 32, 64), enc_num_head=(1, 2, 2, 4),
    enc_patch_size=(64, 64, 64, 64),
    dec_depths=(1, 1, 1), dec_channels=(64, 32, 32), dec_num_head=(4, 2, 2),
    dec_patch_size=(64, 64, 64),
    mlp_ratio=4, qkv_bias=True, drop_path=0.3, shuffle_orders=True, enable_flash=False,
    upcast_attention=True, upcast_softmax=True,
)


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print("wrote", name, os.path.getsize(path) // 1024, "KiB")


def gen_serialization(ref):
    rng = np.random.default_rng(1)
    out = {}
    for depth in (1, 2, 5, 9, 10, 16):
        n = 193
        g = rng.integers(0, 1 << depth, (n, 3)).astype(np.int64)
        g[0] = (1 << depth) - 1  # make sure the max is reached
        g[1] = 0
        b = np.sort(rng.integers(0, 3, n)).astype(np.int64)
        out[f"grid_d{depth}"] = g
        out[f"batch_d{depth}"] = b
        codes = [ref.encode(torch.from_numpy(g), torch.from_numpy(b), depth, o).numpy() for o in ORDERS]
        out[f"code_d{depth}"] = np.stack(codes)
    # Point.serialization incl. randperm consumption (structure.py:47-102)
    g = rng.integers(0, 300, (1500, 3)).astype(np.int64)
    g = np.unique(g, axis=0)
    rng.shuffle(g, axis=0)
    offset = np.array([500, g.shape[0]], dtype=np.int64)
    for shuffle in (False, True):
        torch.manual_seed(1234)
        pt = ref.Point(grid_coord=torch.from_numpy(g), offset=torch.from_numpy(offset))
        pt.serialization(order=ORDERS, shuffle_orders=shuffle)
        tag = "shuf" if shuffle else "noshuf"
        out[f"ps_code_{tag}"] = pt.serialized_code.numpy()
        out[f"ps_order_{tag}"] = pt.serialized_order.numpy()
        out[f"ps_inverse_{tag}"] = pt.serialized_inverse.numpy()
        out[f"ps_depth_{tag}"] = np.int64(pt.serialized_depth)
    torch.manual_seed(1234)
    out["ps_perm"] = torch.randperm(4).numpy()
    out["ps_grid"] = g
    out["ps_offset"] = offset
    save("serialization.npz", **out)


def gen_gridsample(ref):
    out = {}
    cases = {}
    d = synthetic.chunk(20000, L=3.0, H=2.0, seed=3)
    cases["room"] = d["coord"]
    rng = np.random.default_rng(5)
    # boundary values: exact multiples of the grid size in fp32, negatives, duplicates
    k = rng.integers(-200, 200, (4000, 3))
    bnd = (k.astype(np.float32) * np.float32(0.02)).astype(np.float32)
    bnd = np.concatenate([bnd, bnd[:500]], 0)
    cases["boundary"] = bnd
    for name, coord in cases.items():
        gs = ref.GridSample(grid_size=0.02, hash_type="fnv", mode="train", keys=("coord",),
                            return_inverse=True, return_grid_coord=True)
        np.random.seed(7)
        res = gs(dict(coord=coord.copy()))
        out[f"{name}_coord_in"] = coord
        out[f"{name}_coord_out"] = res["coord"]
        out[f"{name}_grid_coord"] = res["grid_coord"].astype(np.int64)
        out[f"{name}_inverse"] = res["inverse"].astype(np.int64)
        # test mode fragments
        gst = ref.GridSample(grid_size=0.02, hash_type="fnv", mode="test", keys=("coord",),
                             return_grid_coord=True)
        parts = gst(dict(coord=coord.copy()))
        out[f"{name}_n_frag"] = np.int64(len(parts))
        out[f"{name}_frag_sizes"] = np.array([p["index"].shape[0] for p in parts], dtype=np.int64)
        # voxel id of every member of fragment 1 (if any) must equal arange(M): store voxel-of-index
        out[f"{name}_frag_last_index"] = parts[-1]["index"].astype(np.int64)
    save("gridsample.npz", **out)


def gen_patch_table(ref):
    out = {}
    att = ref.SerializedAttention.__new__(ref.SerializedAttention)
    cases = [([19], 8), ([8], 8), ([5], 8), ([19, 24, 47], 8), ([16, 17], 8), ([3000], 1024), ([1024, 2049, 2500], 1024)]
    for i, (off, K) in enumerate(cases):
        att.patch_size = K
        pt = ref.Point(offset=torch.tensor(off, dtype=torch.long), batch=torch.zeros(1))
        pad, unpad, cu = ref.SerializedAttention.get_padding_and_inverse(att, pt)
        out[f"c{i}_offset"] = np.array(off, dtype=np.int64)
        out[f"c{i}_K"] = np.int64(K)
        out[f"c{i}_pad"] = pad.numpy()
        out[f"c{i}_unpad"] = unpad.numpy()
        out[f"c{i}_cu"] = cu.numpy()
    out["n_cases"] = np.int64(len(cases))
    save("patch_table.npz", **out)


def gen_pooling(ref):
    import torch.nn as nn
    from functools import partial
    d = synthetic.chunk(6000, L=2.0, H=1.5, seed=11)
    gs = ref.GridSample(grid_size=0.02, hash_type="fnv", mode="train", keys=("coord",), return_grid_coord=True)
    np.random.seed(0)
    res = gs(dict(coord=d["coord"].copy()))
    g = res["grid_coord"].astype(np.int64)
    n = g.shape[0]
    offset = np.array([n // 3, n], dtype=np.int64)
    torch.manual_seed(99)
    pool = ref.SerializedPooling(16, 32, stride=2, norm_layer=partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01),
                                 act_layer=nn.GELU).eval()
    with torch.no_grad():
        pool.norm[0].running_mean.normal_(0, 0.1)
        pool.norm[0].running_var.uniform_(0.5, 1.5)
        pool.norm[0].weight.normal_(1, 0.1)
        pool.norm[0].bias.normal_(0, 0.1)
    feat = torch.randn(n, 16)
    pt = ref.Point(coord=torch.from_numpy(res["coord"]), grid_coord=torch.from_numpy(g), feat=feat,
                   offset=torch.from_numpy(offset))
    torch.manual_seed(4321)
    pt.serialization(order=ORDERS, shuffle_orders=True)
    with torch.no_grad():
        child = pool(pt)
    torch.manual_seed(4321)
    perms = np.stack([torch.randperm(4).numpy() for _ in range(2)])
    out = dict(coord=res["coord"], grid_coord=g, feat=feat.numpy(), offset=offset, perms=perms,
               parent_code=pt.serialized_code.numpy(),
               cluster=child.pooling_inverse.numpy(), code=child.serialized_code.numpy(),
               order=child.serialized_order.numpy(), inverse=child.serialized_inverse.numpy(),
               out_grid_coord=child.grid_coord.numpy(), out_batch=child.batch.numpy(),
               out_coord=child.coord.numpy(), out_feat=child.feat.numpy(),
               depth=np.int64(child.serialized_depth))
    for k, v in pool.state_dict().items():
        out["sd." + k] = v.numpy()
    save("pooling.npz", **out)


def gen_losses(ref):
    torch.manual_seed(5)
    n, dim = 900, 768
    # fp16-representable values so the fixture can store halves exactly
    pred = torch.nn.functional.normalize(torch.randn(n, dim), dim=1).half().float()
    target = torch.nn.functional.normalize(torch.randn(n, dim), dim=1).half().float()
    mask = torch.rand(n) < 0.8
    segment = torch.randint(-1, 6, (n,))
    segment[segment == 4] = 5  # leave a gap in the label set
    segment[400:][segment[400:] == 3] = 0  # class 3 ends up with < 100 valid points -> skipped
    cos = ref.losses.CosineSimilarity()(pred, target, valid_feat_mask=mask)
    l2 = ref.losses.L2Loss()(pred, target, valid_feat_mask=mask)
    # contrastive: capture the half split by replaying the same RNG stream
    crit = ref.losses.AggregatedContrastiveLoss(temperature=0.2, loss_weight=0.025, schedule="all")
    torch.manual_seed(77)
    con = crit(pred, target, valid_feat_mask=mask, segment=segment)
    torch.manual_seed(77)
    valid = (mask > 0) & (segment != -1)
    labels = segment[valid]
    vidx = valid.nonzero(as_tuple=True)[0]
    half = torch.full((n,), -1, dtype=torch.long)
    for lab in torch.unique(labels):
        ind = (labels == lab).nonzero(as_tuple=True)[0]
        if ind.numel() < 100:
            continue
        perm = ind[torch.randperm(ind.size(0))]
        split = perm.size(0) // 2
        half[vidx[perm[:split]]] = 0
        half[vidx[perm[split:]]] = 1
    # lang head with the shipped SigLIP2 ScanNet20 text embeddings
    emb_path = os.path.join(REFERENCE_ROOT, "pointcept/datasets/preprocessing/scannet/meta_data/"
                                            "scannet20_text_embeddings_siglip2.pt")
    text = torch.load(emb_path, map_location="cpu").float()
    logits = torch.mm(pred, text.t())
    probs = torch.sigmoid(logits)
    mx, arg = probs.max(1)
    save("losses.npz", pred=pred.numpy().astype(np.float16), target=target.numpy().astype(np.float16), mask=mask.numpy(),
         segment=segment.numpy(), half=half.numpy(), cos=cos.numpy(), l2=l2.numpy(), con=con.numpy(),
         text=text.numpy(), max_prob=mx.numpy(), argmax=arg.numpy())


def gen_ptv3_small(ref):
    d = synthetic.chunk(9000, L=2.4, H=1.6, seed=21)
    gs = ref.GridSample(grid_size=0.02, hash_type="fnv", mode="train",
                        keys=("coord", "color", "opacity", "quat", "scale"), return_grid_coord=True)
    np.random.seed(0)
    res = gs({k: v.copy() for k, v in d.items()})
    n = res["coord"].shape[0]
    feat = synthetic.feat_from(res)
    offset = np.array([n // 2, n], dtype=np.int64)
    torch.manual_seed(0)
    model = ref.PointTransformerV3(**SMALL_CFG).eval()
    # make BN running stats non-trivial and weights fp16-representable (fixture stores fp16)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
        for p in list(model.parameters()) + list(model.buffers()):
            if p.is_floating_point():
                p.copy_(p.half().float())
    data = dict(coord=torch.from_numpy(res["coord"]), grid_coord=torch.from_numpy(res["grid_coord"].astype(np.int64)),
                feat=torch.from_numpy(feat), offset=torch.from_numpy(offset))
    torch.manual_seed(2024)
    with torch.no_grad():
        out = model(data)
    torch.manual_seed(2024)
    perms = np.stack([torch.randperm(4).numpy() for _ in range(4)])
    arrs = dict(coord=res["coord"], grid_coord=res["grid_coord"].astype(np.int64), feat=feat, offset=offset,
                perms=perms, out_feat=out.feat.numpy(), final_code=out.serialized_code.numpy())
    for k, v in model.state_dict().items():
        arrs["sd." + k] = v.numpy().astype(np.float16) if v.is_floating_point() else v.numpy()
    save("ptv3_small.npz", **arrs)


def gen_ptv3_ssl(ref):
    """PT-v3m1-simdino (point_transformer_v3m1_ssl.py:532-790): mask token, max pooling, (encoder, decoder) outputs."""
    from oracle.ref_shim import load_reference_ssl
    SSL = load_reference_ssl()
    d = synthetic.chunk(7000, L=2.2, H=1.6, seed=23)
    gs = ref.GridSample(grid_size=0.02, hash_type="fnv", mode="train",
                        keys=("coord", "color", "opacity", "quat", "scale"), return_grid_coord=True)
    np.random.seed(1)
    res = gs({k: v.copy() for k, v in d.items()})
    n = res["coord"].shape[0]
    feat = synthetic.feat_from(res)
    offset = np.array([n // 3, n], dtype=np.int64)
    torch.manual_seed(0)
    model = SSL(**SMALL_CFG, do_mask=True, pooling_reduce="max").eval()
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
        for p in list(model.parameters()) + list(model.buffers()):
            if p.is_floating_point():
                p.copy_(p.half().float())
    mask = torch.from_numpy(np.random.default_rng(5).uniform(size=n) < 0.3)
    data = dict(coord=torch.from_numpy(res["coord"]), grid_coord=torch.from_numpy(res["grid_coord"].astype(np.int64)),
                feat=torch.from_numpy(feat), offset=torch.from_numpy(offset))
    torch.manual_seed(2025)
    # the reference Block casts `point.feat` to half unconditionally (ssl.py:330-331): it only runs under AMP
    with torch.no_grad(), torch.autocast(device_type="cpu", dtype=torch.float16):
        enc, dec = model(data, mask=mask, return_dec=True)
    arrs = dict(coord=res["coord"], grid_coord=res["grid_coord"].astype(np.int64), feat=feat, offset=offset,
                mask=mask.numpy(), enc_feat=enc.feat.float().numpy(), enc_offset=enc.offset.numpy(),
                dec_feat=dec.feat.float().numpy())
    for k, v in model.state_dict().items():
        arrs["sd." + k] = v.numpy().astype(np.float16) if v.is_floating_point() else v.numpy()
    save("ptv3_ssl.npz", **arrs)


def gen_spherecrop(ref):
    """SphereCrop (transform.py:1419-1535), modes center and random, under a fixed numpy seed."""
    out = {}
    d = synthetic.chunk(30000, L=4.0, H=2.5, seed=11)
    coord = d["coord"]
    seg = (np.arange(coord.shape[0]) % 37).astype(np.int64)
    out["coord_in"] = coord
    out["segment_in"] = seg
    for mode in ("center", "random"):
        np.random.seed(21)
        sc = ref.SphereCrop(point_max=7000, mode=mode)
        res = sc(dict(coord=coord.copy(), segment=seg.copy(), color=d["color"].copy()))
        out[f"{mode}_coord"] = res["coord"]
        out[f"{mode}_segment"] = res["segment"]
        out[f"{mode}_color"] = res["color"]
    np.random.seed(22)
    sc = ref.SphereCrop(sample_rate=0.25, mode="random")
    res = sc(dict(coord=coord.copy(), segment=seg.copy()))
    out["rate_coord"] = res["coord"]
    out["rate_segment"] = res["segment"]
    save("spherecrop.npz", **out)


def gen_spherecrop_all(ref):
    """SphereCrop(mode="all") (transform.py:1439-1503): the list of crops under a fixed numpy seed."""
    import contextlib, io
    d = synthetic.chunk(12000, L=3.0, H=2.0, seed=13)
    coord = d["coord"]
    np.random.seed(31)
    sc = ref.SphereCrop(point_max=4000, mode="all")
    with contextlib.redirect_stdout(io.StringIO()):  # (the reference prints the dict keys of every crop)
        parts = sc(dict(coord=coord.copy(), color=d["color"].copy(), opacity=d["opacity"].copy()))
    out = dict(coord_in=coord, color_in=d["color"], opacity_in=d["opacity"], n_parts=np.array(len(parts)))
    for i, p in enumerate(parts):
        out[f"p{i}_index"] = p["index"]
        out[f"p{i}_weight"] = p["weight"]
        out[f"p{i}_coord"] = p["coord"]
        out[f"p{i}_color"] = p["color"]
    save("spherecrop_all.npz", **out)


if __name__ == "__main__":
    ref = load_reference()
    which = sys.argv[1:] or ["serialization", "gridsample", "patch_table", "pooling", "losses", "ptv3_small", "ptv3_ssl", "spherecrop", "spherecrop_all"]
    for w in which:
        globals()["gen_" + w](ref)
