"""GPU parity tests (through the C-ABI) of the integer / index stages: bit-exact against the oracle and
the golden vectors generated from the reference."""
import numpy as np
import pytest
import torch

from oracle import attention as oattn
from oracle import gridsample as ogs
from oracle import pooling as opool
from oracle import serialization as oser
from oracle import subm_conv as oconv
from scenesplat_b200 import synthetic

pytestmark = pytest.mark.gpu
ORDERS = oser.ORDERS


def dev(a):
    return torch.as_tensor(a).cuda()


@pytest.mark.parametrize("depth", [1, 2, 5, 9, 10, 16])
def test_serialize_golden_codes(golden, depth):
    from scenesplat_b200 import ops
    g = golden("serialization.npz")
    grid, batch, ref = g[f"grid_d{depth}"], g[f"batch_d{depth}"], g[f"code_d{depth}"]
    offset = np.cumsum(np.bincount(batch, minlength=3))
    b, code, order, inv = ops.serialize(dev(grid), dev(offset), depth, ORDERS)
    np.testing.assert_array_equal(code.cpu().numpy(), ref)
    np.testing.assert_array_equal(b.cpu().numpy(), batch)
    _, o_order, o_inv, _ = oser.serialization(grid, batch, 3, ORDERS, depth=depth)
    np.testing.assert_array_equal(order.cpu().numpy(), o_order)
    np.testing.assert_array_equal(inv.cpu().numpy(), o_inv)


@pytest.mark.parametrize("shuffle", [False, True])
def test_point_serialization_golden(golden, shuffle):
    from scenesplat_b200 import ops
    g = golden("serialization.npz")
    tag = "shuf" if shuffle else "noshuf"
    grid, offset = g["ps_grid"], g["ps_offset"]
    depth = ops.coord_depth(dev(grid))
    assert depth == int(g[f"ps_depth_{tag}"])
    orders = [ORDERS[i] for i in g["ps_perm"]] if shuffle else ORDERS
    _, code, order, inv = ops.serialize(dev(grid), dev(offset), depth, orders)
    np.testing.assert_array_equal(code.cpu().numpy(), g[f"ps_code_{tag}"])
    np.testing.assert_array_equal(order.cpu().numpy(), g[f"ps_order_{tag}"])
    np.testing.assert_array_equal(inv.cpu().numpy(), g[f"ps_inverse_{tag}"])


@pytest.mark.parametrize("n,nb,int32", [(1, 1, False), (255, 1, True), (2049, 3, False), (70001, 5, True),
                                        (300000, 1, False)])
def test_serialize_vs_oracle_sizes(n, nb, int32):
    from scenesplat_b200 import ops
    rng = np.random.default_rng(n)
    grid = rng.integers(0, 700, (n, 3))
    if n > 10:  # duplicates exercise the stable tie rule
        grid[n // 2] = grid[0]
    cuts = np.sort(rng.choice(np.arange(1, n), nb - 1, replace=False)) if nb > 1 else np.array([], dtype=np.int64)
    offset = np.concatenate([cuts, [n]]).astype(np.int64)
    batch = oser.offset2batch(offset)
    gd = dev(grid.astype(np.int32 if int32 else np.int64))
    depth = ops.coord_depth(gd)
    assert depth == oser.serialization_depth(grid)
    b, code, order, inv = ops.serialize(gd, dev(offset), depth, ORDERS)
    o_code, o_order, o_inv, _ = oser.serialization(grid, batch, nb, ORDERS)
    np.testing.assert_array_equal(b.cpu().numpy(), batch)
    np.testing.assert_array_equal(code.cpu().numpy(), o_code)
    np.testing.assert_array_equal(order.cpu().numpy(), o_order)
    np.testing.assert_array_equal(inv.cpu().numpy(), o_inv)
    # size-independent properties: order is a permutation sorting the code, inverse inverts it
    c = code.gather(1, order)
    assert bool((c[:, 1:] >= c[:, :-1]).all())
    ar = torch.arange(n, device="cuda").expand(4, n)
    assert bool((inv.gather(1, order) == ar).all())


def test_serialize_empty():
    from scenesplat_b200 import ops
    g = torch.zeros((0, 3), dtype=torch.int64, device="cuda")
    b, code, order, inv = ops.serialize(g, dev(np.array([0])), 1, ORDERS)
    assert code.shape == (4, 0) and order.shape == (4, 0)


@pytest.mark.parametrize("name", ["room", "boundary"])
def test_gridsample_golden(golden, name):
    from scenesplat_b200 import ops
    g = golden("gridsample.npz")
    coord = g[f"{name}_coord_in"]
    ix = ops.gridsample_index(dev(coord), 0.02)
    o = ogs.grid_sample_index(coord, 0.02)
    assert ix["m"] == o["count"].shape[0]
    np.testing.assert_array_equal(ix["inverse"].cpu().numpy(), g[f"{name}_inverse"])
    np.testing.assert_array_equal(ix["idx_sort"].cpu().numpy(), o["idx_sort"])
    np.testing.assert_array_equal(ix["start"][: ix["m"]].cpu().numpy(), o["start"])
    np.testing.assert_array_equal(ix["min_coord"].cpu().numpy(), o["min_coord"])
    rng = np.random.default_rng(0)
    rnd = rng.integers(0, int(o["count"].max()), o["count"].size)
    idx_u, gc, cnt = ops.gridsample_select(ix, dev(rnd), want_count=True)
    ref = ogs.grid_sample_train(coord, 0.02, rnd)
    np.testing.assert_array_equal(idx_u.cpu().numpy(), ref["idx_unique"])
    np.testing.assert_array_equal(gc.cpu().numpy(), g[f"{name}_grid_coord"])
    np.testing.assert_array_equal(cnt.cpu().numpy(), ref["count"])
    # attribute gather
    out = ops.gather_rows(dev(coord), idx_u)
    np.testing.assert_array_equal(out.cpu().numpy(), coord[ref["idx_unique"]])
    # test-mode fragments
    parts, _ = ogs.grid_sample_test(coord, 0.02)
    for i in (0, len(parts) - 1):
        idx_f, _, _ = ops.gridsample_select(ix, None, frag=i, want_grid_coord=False)
        np.testing.assert_array_equal(idx_f.cpu().numpy(), parts[i])


def test_gridsample_ravel_and_large():
    from scenesplat_b200 import ops
    d = synthetic.chunk(200000, seed=9)
    for ht in ("fnv", "ravel"):
        ix = ops.gridsample_index(dev(d["coord"]), 0.02, ht)
        o = ogs.grid_sample_index(d["coord"], 0.02, ht)
        assert ix["m"] == o["count"].shape[0]
        np.testing.assert_array_equal(ix["inverse"].cpu().numpy(), o["inverse"])
        np.testing.assert_array_equal(ix["idx_sort"].cpu().numpy(), o["idx_sort"])
    # property: inverse is constant per voxel and the number of distinct values is m
    inv = ix["inverse"]
    assert int(inv.max()) + 1 == ix["m"]


def _parent(n_raw=30000, seed=4, nb=2):
    d = synthetic.chunk(n_raw, L=3.0, H=2.0, seed=seed)
    res = ogs.grid_sample_train(d["coord"], 0.02)
    g = res["grid_coord"]
    n = g.shape[0]
    offset = np.array([n // 3, n] if nb == 2 else [n], dtype=np.int64)
    return d["coord"][res["idx_unique"]], g, offset


@pytest.mark.parametrize("perm", [[0, 1, 2, 3], [2, 0, 3, 1], [3, 2, 1, 0]])
def test_pool_index_vs_oracle(perm):
    from scenesplat_b200 import ops
    coord, g, offset = _parent()
    batch = oser.offset2batch(offset)
    code, order, inv, depth = oser.serialization(g, batch, len(offset), ORDERS, perm=[1, 3, 0, 2])
    ref = opool.pool_index(code, 1, perm=perm)
    out = ops.pool_index(dev(code), dev(order), dev(g), dev(batch), 1, perm)
    assert out["m"] == ref["counts"].shape[0]
    np.testing.assert_array_equal(out["cluster"].cpu().numpy(), ref["cluster"])
    np.testing.assert_array_equal(out["code"].cpu().numpy(), ref["code"])
    np.testing.assert_array_equal(out["order"].cpu().numpy(), ref["order"])
    np.testing.assert_array_equal(out["inverse"].cpu().numpy(), ref["inverse"])
    np.testing.assert_array_equal(np.diff(out["seg_start"].cpu().numpy()), ref["counts"])
    gc, b = opool.pooled_attrs(g, batch, ref["head"], 1)
    np.testing.assert_array_equal(out["grid_coord"].cpu().numpy(), gc)
    np.testing.assert_array_equal(out["batch"].cpu().numpy(), b)
    # segment mean (fp32) of features and coords
    rng = np.random.default_rng(0)
    feat = rng.normal(size=(g.shape[0], 64)).astype(np.float32)
    want = opool.segment_csr(feat, ref["indices"], ref["idx_ptr"], "mean")
    got = ops.segment_reduce(dev(feat), dev(order[0]), out["seg_start"], "mean")
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-6)
    wantc = opool.segment_csr(coord, ref["indices"], ref["idx_ptr"], "mean")
    gotc = ops.segment_reduce(dev(coord), dev(order[0]), out["seg_start"], "mean")
    np.testing.assert_allclose(gotc.cpu().numpy(), wantc, rtol=1e-5, atol=1e-6)
    for red in ("sum", "max", "min"):
        w = opool.segment_csr(feat, ref["indices"], ref["idx_ptr"], red)
        gg = ops.segment_reduce(dev(feat), dev(order[0]), out["seg_start"], red)
        np.testing.assert_allclose(gg.cpu().numpy(), w, rtol=1e-5, atol=1e-5)
    # unpool gather-add
    child = rng.normal(size=(out["m"], 64)).astype(np.float32)
    u, _ = ops.unpool_gather_add(dev(feat), dev(child), out["cluster"])
    np.testing.assert_allclose(u.cpu().numpy(), opool.unpool_gather_add(feat, child, ref["cluster"]), rtol=1e-6)


def test_pool_golden(golden):
    from scenesplat_b200 import ops
    g = golden("pooling.npz")
    offset = g["offset"]
    perms = g["perms"]
    depth = ops.coord_depth(dev(g["grid_coord"]))
    b, code, order, inv = ops.serialize(dev(g["grid_coord"]), dev(offset), depth, [ORDERS[i] for i in perms[0]])
    np.testing.assert_array_equal(code.cpu().numpy(), g["parent_code"])
    out = ops.pool_index(code, order, dev(g["grid_coord"]), b, 1, list(perms[1]))
    np.testing.assert_array_equal(out["cluster"].cpu().numpy(), g["cluster"])
    np.testing.assert_array_equal(out["code"].cpu().numpy(), g["code"])
    np.testing.assert_array_equal(out["order"].cpu().numpy(), g["order"])
    np.testing.assert_array_equal(out["inverse"].cpu().numpy(), g["inverse"])
    np.testing.assert_array_equal(out["grid_coord"].cpu().numpy(), g["out_grid_coord"])
    np.testing.assert_array_equal(out["batch"].cpu().numpy(), g["out_batch"])
    # fused mean -> BN(eval) -> GELU epilogue against the reference module's output
    sd = {k[3:]: torch.from_numpy(g[k]).cuda() for k in g.files if k.startswith("sd.")}
    proj = torch.nn.functional.linear(dev(g["feat"]), sd["proj.weight"], sd["proj.bias"])
    scale = sd["norm.0.weight"] / torch.sqrt(sd["norm.0.running_var"] + 1e-3)
    shift = sd["norm.0.bias"] - sd["norm.0.running_mean"] * scale
    feat = ops.segment_reduce(proj, order[0].contiguous(), out["seg_start"], "mean", scale.contiguous(),
                              shift.contiguous(), act=1)
    np.testing.assert_allclose(feat.cpu().numpy(), g["out_feat"], rtol=2e-4, atol=2e-5)
    coord = ops.segment_reduce(dev(g["coord"]), order[0].contiguous(), out["seg_start"], "mean")
    np.testing.assert_allclose(coord.cpu().numpy(), g["out_coord"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("k,row", [(3, 0), (3, 2), (5, 1), (3, 3)])
def test_kernel_map_vs_oracle(k, row):
    from scenesplat_b200 import ops
    _, g, offset = _parent(20000, seed=6)
    batch = oser.offset2batch(offset)
    code, order, inv, depth = oser.serialization(g, batch, len(offset), ORDERS)
    nbr, cnt = ops.kmap_build(dev(g), dev(batch), dev(code[row]), dev(order[row]), depth, row, k)
    ref = oconv.kernel_map(g, batch, k)
    np.testing.assert_array_equal(nbr.cpu().numpy().T, ref)
    np.testing.assert_array_equal(cnt.cpu().numpy(), (ref >= 0).sum(0))
    pairs = ops.kmap_pairs(nbr, dev(order[row]), k, cnt.cpu().numpy())
    ypos = pairs["ypos"].cpu().numpy().T  # [n, k3]
    pin = pairs["pair_in"].cpu().numpy()
    assert ((ypos >= 0) == (ref >= 0)).all()
    sel = ref >= 0
    np.testing.assert_array_equal(pin[ypos[sel]], ref[sel])
    assert np.unique(ypos[sel]).size == sel.sum()
    assert pairs["p_pad"] % 128 == 0


def test_kernel_map_subset_equals_search():
    """The 3^3 map taken as a row subset of the 5^3 map (ss_kmap_subset: what PointTransformerV3.plan_indices does at level 0
    behind the stem's map) is the map a search of its own produces, counts included; a duplicated voxel resolves to the
    smallest index as the lower-bound search of rounds 1 - 2 did."""
    from scenesplat_b200 import ops
    _, g, offset = _parent(15000, seed=9)
    batch = oser.offset2batch(offset)
    code, order, inv, depth = oser.serialization(g, batch, len(offset), ORDERS)
    args = (dev(g), dev(batch), dev(code[0]), dev(order[0]), depth, 0)
    nbr5, cnt5 = ops.kmap_build(*args, 5)
    nbr3, cnt3 = ops.kmap_build(*args, 3)
    sub, csub = ops.kmap_subset(nbr5, cnt5, 5, 3)
    assert torch.equal(sub, nbr3) and torch.equal(csub, cnt3)
    ref = oconv.kernel_map(g, batch, 3)
    np.testing.assert_array_equal(sub.cpu().numpy().T, ref)
    # a duplicated voxel (legal input, not produced by GridSample): the LOOKED-UP taps (t < 13) of every other voxel
    # resolve to the smaller index (the mirrored taps are written by both copies, as in rounds 1 - 2)
    _, g1, off1 = _parent(6000, seed=10, nb=1)
    g2 = np.concatenate([g1, g1[:1]], 0)
    b2 = np.zeros(g2.shape[0], dtype=np.int64)
    code2, order2, _, depth2 = oser.serialization(g2, b2, 1, ORDERS)
    nbr_d, _ = ops.kmap_build(dev(g2), dev(b2), dev(code2[0]), dev(order2[0]), depth2, 0, 3)
    nd = nbr_d.cpu().numpy()
    assert not (nd[:13, :-1] == g2.shape[0] - 1).any()
    np.testing.assert_array_equal(nd[:13, -1], nd[:13, 0])  # the copy sees voxel 0's neighbours


def test_patch_table_vs_golden(golden):
    from scenesplat_b200 import ops
    g = golden("patch_table.npz")
    for i in range(int(g["n_cases"])):
        offset, K = g[f"c{i}_offset"], int(g[f"c{i}_K"])
        n = int(offset[-1])
        table = ops.patch_table(dev(offset), K, n).cpu().numpy()
        pad, unpad, cu = g[f"c{i}_pad"], g[f"c{i}_unpad"], g[f"c{i}_cu"]
        # rebuild pad / unpad / cu_seqlens from the table and compare with the reference's arrays
        rows = [t for t in table if t[1] > t[0]]
        assert len(rows) == len(cu) - 1
        seq_pad, seq_unpad = [], np.zeros(n, dtype=np.int64)
        pos = 0
        for qb, qe, kb, ke in rows:
            own = list(range(qb, qe))
            borrowed = [j for j in range(kb, ke) if j < qb]
            for t, j in enumerate(own):
                seq_unpad[j] = pos + t
            seq_pad += own + borrowed
            pos += (ke - kb)
        np.testing.assert_array_equal(np.array(seq_pad), pad)
        np.testing.assert_array_equal(seq_unpad, unpad)
        np.testing.assert_array_equal(np.cumsum([0] + [r[3] - r[2] for r in rows]), cu)


@pytest.mark.parametrize("n,k,with_mask,with_query", [(20000, 25, True, False), (5000, 25, False, True), (300, 7, True, True),
                                                      (40, 25, False, False)])
def test_neighbor_voting_and_confusion(n, k, with_mask, with_query):
    """GPU k-NN majority vote (csrc/voting.cu) vs scipy cKDTree + numpy vote (oracle/voting.py, restating
    pointcept/utils/misc.py:17-95): bit-exact labels (fp64 distances on both sides; real-valued synthetic coordinates
    have no ties at the k-th neighbour).  Confusion-matrix update vs the reference's per-point loop: bit-exact."""
    import scenesplat_b200 as S
    from oracle import voting as ovote
    from scenesplat_b200 import synthetic
    rng = np.random.default_rng(n + k)
    coords = synthetic.room(n, L=4.0, H=2.5, seed=3)
    num_classes, ignore = 20, -1
    # spatially coherent labels with noise, some ignored
    pred = ((coords[:, 0] * 1.7 + coords[:, 1] * 2.3).astype(np.int64) % num_classes)
    noise = rng.random(n) < 0.3
    pred[noise] = rng.integers(0, num_classes, noise.sum())
    pred[rng.random(n) < 0.1] = ignore
    mask = (rng.random(n) < 0.8) if with_mask else None
    query = (coords[rng.integers(0, n, n // 2)] + rng.normal(0, 0.05, (n // 2, 3))).astype(np.float32) if with_query else None
    if with_query:
        query[:5] += 30.0  # far outside the grid
    want = ovote.neighbor_voting(coords, pred, min(k, int(mask.sum()) if mask is not None else n), ignore, num_classes, mask, query)
    got = S.neighbor_voting(dev(coords), dev(pred), k, ignore, num_classes,
                            dev(mask) if mask is not None else None, dev(query) if query is not None else None)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    # confusion matrix
    gt = rng.integers(0, num_classes, want.shape[0])
    conf, fn = np.zeros((num_classes, num_classes), np.int64), np.zeros(num_classes, np.int64)
    ovote.confusion_update(gt, want, num_classes, ignore, conf, fn)
    conf_d = torch.zeros((num_classes, num_classes), dtype=torch.int64, device="cuda")
    fn_d = torch.zeros(num_classes, dtype=torch.int64, device="cuda")
    S.confusion_update(dev(gt), got, num_classes, ignore, conf_d, fn_d)
    np.testing.assert_array_equal(conf_d.cpu().numpy(), conf)
    np.testing.assert_array_equal(fn_d.cpu().numpy(), fn)


def test_sphere_crop_golden(golden):
    """GPU SphereCrop vs the reference's own output (tests/golden/spherecrop.npz, generated from the unmodified
    reference under fixed numpy seeds): the same rows at the same (ascending, bit-identical fp32) distances.  Rows at
    EXACTLY equal distance may be permuted (numpy's argsort is an unstable introsort, the GPU sort breaks ties by
    index), so both sides are canonicalised inside tie groups before the bit-exact comparison."""
    import scenesplat_b200 as S
    g = golden("spherecrop.npz")
    from scenesplat_b200 import synthetic
    color = synthetic.chunk(30000, L=4.0, H=2.5, seed=11)["color"]

    def canon(res, center):
        d = np.sum(np.square(res["coord"] - center), 1)
        assert np.all(np.diff(d) >= 0)  # ascending distance
        c = res["coord"]
        perm = np.lexsort((c[:, 2], c[:, 1], c[:, 0], d))
        return d, {k: v[perm] for k, v in res.items()}

    n = g["coord_in"].shape[0]
    for mode in ("center", "random"):
        np.random.seed(21)
        ci = np.random.randint(n) if mode == "random" else n // 2
        center = g["coord_in"][ci]
        np.random.seed(21)
        sc = S.SphereCrop(point_max=7000, mode=mode)
        res = sc(dict(coord=g["coord_in"].copy(), segment=g["segment_in"].copy(), color=color.copy()))
        want = dict(coord=g[f"{mode}_coord"], segment=g[f"{mode}_segment"], color=g[f"{mode}_color"])
        d_got, got_c = canon(res, center)
        d_want, want_c = canon(want, center)
        np.testing.assert_array_equal(d_got, d_want)
        for k in want:
            np.testing.assert_array_equal(got_c[k], want_c[k])
    np.random.seed(22)
    ci = np.random.randint(n)
    np.random.seed(22)
    res = S.SphereCrop(sample_rate=0.25, mode="random")(dict(coord=g["coord_in"].copy(), segment=g["segment_in"].copy()))
    d_got, got_c = canon(res, g["coord_in"][ci])
    d_want, want_c = canon(dict(coord=g["rate_coord"], segment=g["rate_segment"]), g["coord_in"][ci])
    np.testing.assert_array_equal(d_got, d_want)
    np.testing.assert_array_equal(got_c["coord"], want_c["coord"])
    np.testing.assert_array_equal(got_c["segment"], want_c["segment"])
    # torch tensors in -> torch tensors (on the GPU) out; fewer points than point_max -> untouched
    d = dict(coord=dev(g["coord_in"][:100]))
    assert S.SphereCrop(point_max=1000)(d)["coord"].shape[0] == 100


def test_packed_scene_one_copy_to_device(tmp_path):
    """scene_io: the whole scene arrives on the GPU with ONE copy from ONE pinned buffer; every attribute is a view of
    that device buffer and holds exactly the bytes the reference's get_data would produce (oracle/scene_io.py)."""
    import os
    from oracle import scene_io as oio
    from scenesplat_b200 import scene_io as sio
    from tests.test_cpu_host import _write_scene_folder
    folder = str(tmp_path / "scene0001_00")
    _write_scene_folder(folder, n=20000, seed=2)
    want = oio.get_data(folder, is_train=True)
    path = str(tmp_path / "scene0001_00.sspk")
    sio.pack_scene(folder, path)
    sc = sio.load_scene(path, keys=tuple(want))
    assert sc.host.is_pinned()
    dev_arrays = sc.to_device("cuda")
    torch.cuda.synchronize()
    base = {v.untyped_storage().data_ptr() for v in dev_arrays.values()}
    assert len(base) == 1                                   # one device allocation behind all attributes
    for k, v in want.items():
        got = dev_arrays[k]
        assert got.is_cuda and tuple(got.shape) == v.shape, k
        np.testing.assert_array_equal(got.cpu().numpy(), v)
    assert dev_arrays["lang_feat"].dtype == torch.float16   # stays fp16 until a kernel reads it


def test_sphere_crop_all_golden(golden):
    """SphereCrop(mode="all") (test-time multi-crop, transform.py:1439-1503) vs the reference's own list of crops under
    the same numpy seed: same number of crops, same member sets, bit-identical fp32 `weight` (squared distance to the
    seed), same gathered attributes.  Members at exactly equal distance may be permuted (numpy's unstable argsort), so
    each crop is compared as a set keyed by `index`."""
    import scenesplat_b200 as S
    g = golden("spherecrop_all.npz")
    np.random.seed(31)
    parts = S.SphereCrop(point_max=4000, mode="all")(dict(coord=g["coord_in"].copy(), color=g["color_in"].copy(),
                                                         opacity=g["opacity_in"].copy()))
    assert len(parts) == int(g["n_parts"])
    covered = np.zeros(g["coord_in"].shape[0], dtype=bool)
    for i, p in enumerate(parts):
        assert set(p.keys()) == {"coord", "color", "opacity", "weight", "index"}
        a, b = np.argsort(p["index"], kind="stable"), np.argsort(g[f"p{i}_index"], kind="stable")
        np.testing.assert_array_equal(p["index"][a], g[f"p{i}_index"][b])
        np.testing.assert_array_equal(p["weight"][a], g[f"p{i}_weight"][b])
        np.testing.assert_array_equal(p["coord"][a], g[f"p{i}_coord"][b])
        np.testing.assert_array_equal(p["color"][a], g[f"p{i}_color"][b])
        assert np.all(np.diff(p["weight"]) >= 0)
        covered[p["index"]] = True
    assert covered.all()
    # fewer points than point_max: one part holding everything, zero weights
    small = S.SphereCrop(point_max=10 ** 6, mode="all")(dict(coord=g["coord_in"][:100].copy()))
    assert len(small) == 1 and small[0]["weight"].shape == (100,) and not small[0]["weight"].any()
    np.testing.assert_array_equal(small[0]["index"], np.arange(100))


def test_lang_feat_stays_fp16_end_to_end():
    """`lang_feat` is stored as fp16 on disk (datasets/scannetgs.py); the reference's ToTensor blows it up to fp32
    (transform.py:390-391).  Here it stays fp16 through GridSample, SphereCrop and into the fused cosine / L2 losses
    (forward and adjoint): same loss and same gradient as with an fp32 copy, at a quarter of the bytes."""
    import scenesplat_b200 as S
    from scenesplat_b200 import synthetic, training
    d = synthetic.chunk(20000, L=2.4, H=1.6, seed=9, lang_dim=768)
    assert d["lang_feat"].dtype == np.float16
    gs = S.GridSample(grid_size=0.02, hash_type="fnv", mode="train",
                      keys=("coord", "color", "opacity", "quat", "scale", "lang_feat", "valid_feat_mask", "segment"),
                      return_grid_coord=True)
    np.random.seed(0)
    out = gs({k: torch.from_numpy(v) for k, v in d.items()})
    np.random.seed(1)
    out = S.SphereCrop(point_max=9000, mode="random")(out)
    assert out["lang_feat"].dtype == torch.float16 and out["lang_feat"].is_cuda and out["lang_feat"].shape == (9000, 768)
    torch.manual_seed(0)
    res = {}
    for name, tgt in (("fp16", out["lang_feat"]), ("fp32", out["lang_feat"].float())):
        pred = torch.randn(9000, 768, device="cuda", generator=torch.Generator("cuda").manual_seed(4)).requires_grad_(True)
        loss = training.cosine_loss(pred, tgt, out["valid_feat_mask"], 1.0) + training.l2_loss(pred, tgt, out["valid_feat_mask"], 1.0)
        loss.backward()
        res[name] = (loss.item(), pred.grad.clone())
    assert abs(res["fp16"][0] - res["fp32"][0]) < 1e-6 * abs(res["fp32"][0])
    assert torch.equal(res["fp16"][1], res["fp32"][1])
    acc16 = S.CosineSimilarity()(torch.randn(9000, 768, device="cuda"), out["lang_feat"], out["valid_feat_mask"])
    assert torch.isfinite(acc16)
