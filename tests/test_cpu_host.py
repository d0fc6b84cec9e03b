"""CPU tests of the host side: the C-ABI library loads (no GPU needed) and exports exactly the symbols
include/scenesplat_b200.h declares; registries / Point dict / module tree mirror the reference; the
product path refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "scenesplat_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ss_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    from scenesplat_b200 import _lib, build
    path = build.build()
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.SIGNATURES) == declared  # the ctypes table mirrors the header one to one
    assert _lib.load().ss_version().startswith(b"scenesplat_b200")
    assert _lib.load().ss_launch_count() == 0  # nothing was launched: loading is not computing


def test_header_cites_reference_interfaces():
    text = open(os.path.join(ROOT, "include", "scenesplat_b200.h")).read()
    for ref in ("structure.py:47-102", "transform.py:1211-1330", "point_transformer_v3m1_base.py:371-444",
                "point_transformer_v3m1_base.py:114-222", "losses/misc.py:247-295", "evaluator.py:793-800"):
        assert ref in text, ref


def test_no_cpu_fallback():
    from scenesplat_b200 import _lib, ops
    with pytest.raises(_lib.CudaKernelError):
        ops.serialize(torch.zeros((4, 3), dtype=torch.int64), torch.tensor([4]), 3, ["z"])


def test_bf16_shadow_registry_and_weight_folds():
    """Host logic that needs no GPU: the registry that hands the zero-shot head the bf16 rows written beside the fp32
    features (dropped on in-place modification and when the tensor dies), and the Linear + BatchNorm1d(eval) fold of
    SerializedUnpooling (composed in fp32, refreshed when a parameter or a running statistic changes)."""
    import gc
    from scenesplat_b200 import ops, ptv3
    feat = torch.randn(5, 8)
    sh = feat.bfloat16()
    ops._set_bf16_shadow(feat, sh)
    assert ops.bf16_shadow(feat) is sh and ops._head_operand(feat, False) is sh
    assert ops.bf16_shadow(torch.randn(5, 8)) is None
    feat.add_(1.0)
    assert ops.bf16_shadow(feat) is None
    other = torch.randn(3, 8)
    ops._set_bf16_shadow(other, other.bfloat16())
    key = id(other)
    del other
    gc.collect()
    assert key not in ops._BF16_SHADOW

    torch.manual_seed(0)
    lin, bn = torch.nn.Linear(16, 32), torch.nn.BatchNorm1d(32, eps=1e-3)
    bn.running_mean.normal_()
    bn.running_var.uniform_(0.5, 2.0)
    bn.weight.data.normal_()
    bn.bias.data.normal_()
    bn.eval()
    x = torch.randn(7, 16)
    w, b = ptv3.linear_bn_params(lin, bn)
    assert w.dtype == torch.bfloat16 and b.dtype == torch.float32
    want = bn(lin(x)).detach()
    got = x @ w.float().t() + b
    assert float((got - want).abs().max()) < 3e-2 * float(want.abs().max())  # bf16 weights
    assert ptv3.linear_bn_params(lin, bn)[0] is w  # cached
    bn.running_mean.add_(1.0)
    w2, b2 = ptv3.linear_bn_params(lin, bn)
    assert w2 is not w and float((x @ w2.float().t() + b2 - bn(lin(x)).detach()).abs().max()) < 3e-2 * float(want.abs().max())
    assert ptv3.linear_bn_params(lin, None)[0].shape == (32, 16)


def test_state_dict_matches_reference_golden_keys(golden):
    import scenesplat_b200 as S
    from tests.golden.make_golden import SMALL_CFG
    g = golden("ptv3_small.npz")
    ref_keys = sorted(k[3:] for k in g.files if k.startswith("sd."))
    model = S.PointTransformerV3(**SMALL_CFG)
    sd = model.state_dict()
    assert sorted(sd.keys()) == ref_keys
    for k in ref_keys:
        assert tuple(sd[k].shape) == tuple(g["sd." + k].shape), k


def test_lang_config_module_tree():
    import scenesplat_b200 as S
    from oracle.ref_shim import LANG_BACKBONE_CFG
    m = S.build_model(dict(type="LangPretrainer", backbone=dict(type="PT-v3m1", **LANG_BACKBONE_CFG),
                           criteria=[dict(type="CosineSimilarity", reduction="mean", loss_weight=1.0),
                                     dict(type="L2Loss", reduction="mean", loss_weight=1.0),
                                     dict(type="AggregatedContrastiveLoss", temperature=0.2, reduction="mean",
                                          loss_weight=0.025, schedule="all")]))
    sd = m.state_dict()
    assert len(sd) == 393
    assert sum(p.numel() for p in m.parameters()) == 91712800
    assert tuple(sd["backbone.embedding.stem.conv.weight"].shape) == (32, 5, 5, 5, 11)
    assert tuple(sd["backbone.dec.dec0.block0.cpe.0.weight"].shape) == (768, 3, 3, 3, 768)
    assert "backbone.enc.enc1.down.norm.0.running_mean" in sd and "backbone.dec.dec0.up.proj_skip.1.weight" in sd
    assert len(m.criteria.criteria) == 3


def test_registry_and_point_dict():
    import scenesplat_b200 as S
    from scenesplat_b200.registry import Registry
    assert S.MODELS.get("PT-v3m1") is S.PointTransformerV3
    assert S.TRANSFORMS.get("GridSample") is S.GridSample
    r = Registry("x")
    r.register_module("A", module=int)
    with pytest.raises(KeyError):
        r.register_module("A", module=float)
    r.register_module("A", module=float, force=True)
    p = S.Point(coord=torch.zeros(5, 3), offset=torch.tensor([2, 5]))
    assert p.batch.tolist() == [0, 0, 1, 1, 1] and p["batch"] is p.batch
    q = S.Point(batch=torch.tensor([0, 0, 1]))
    assert q.offset.tolist() == [2, 3]
    p.feat = torch.ones(5, 1)
    assert "feat" in p.keys()
    with pytest.raises(AttributeError):
        _ = p.nonexistent


def test_training_and_unsupported_paths_raise_loudly():
    import scenesplat_b200 as S
    blk = S.Block(channels=32, num_heads=2, patch_size=64, enable_flash=True, upcast_attention=False, upcast_softmax=False)
    blk.train()
    with pytest.raises(NotImplementedError):
        blk(S.Point(feat=torch.zeros(4, 32), offset=torch.tensor([4])))
    with pytest.raises(NotImplementedError):
        S.PointTransformerV3(pdnorm_bn=True)
    with pytest.raises(NotImplementedError):
        S.GridSample(importance_sample_key="scale_max")


def test_sharding_tables():
    from scenesplat_b200.sharding import assign_chunks, chunk_ranges, job_throughput
    sizes = [300, 100, 250, 50, 400, 120, 80]
    for policy in ("round_robin", "lpt"):
        for w in (1, 2, 4, 8):
            parts = assign_chunks(sizes, w, policy)
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(len(sizes)))
    lpt = assign_chunks(sizes, 2, "lpt")
    loads = [sum(sizes[i] for i in p) for p in lpt]
    assert max(loads) - min(loads) <= max(sizes)
    assert chunk_ranges(10, 4) == [(0, 4), (4, 8), (8, 10)]  # default.py:134-139
    assert job_throughput([10, 10], [1000.0, 2000.0]) == 10.0


def test_synthetic_chunk_statistics():
    from scenesplat_b200 import synthetic
    d = synthetic.chunk(5000, seed=1, lang_dim=8)
    assert d["coord"].dtype == np.float32 and d["opacity"].shape == (5000, 1)
    assert np.all(d["quat"][:, 0] >= 0) and np.allclose(np.linalg.norm(d["quat"], axis=1), 1, atol=1e-5)
    assert d["scale"].max() <= 1.5 and d["color"].min() >= -1 and d["color"].max() <= 1
    assert synthetic.feat_from(d).shape == (5000, 11)


def test_compat_install_exposes_reference_import_names():
    """The names point_transformer_v3m1_base.py:13-24 imports resolve to the stand-ins after install(force=True)."""
    import importlib
    import sys
    from scenesplat_b200 import compat
    saved = {k: sys.modules.get(k) for k in ("torch_scatter", "flash_attn", "spconv", "spconv.pytorch", "spconv.pytorch.modules")}
    try:
        compat.install(force=True)
        spconv = importlib.import_module("spconv.pytorch")
        from spconv.pytorch.modules import is_spconv_module  # noqa: F401
        import flash_attn
        import torch_scatter
        assert callable(torch_scatter.segment_csr) and callable(flash_attn.flash_attn_varlen_qkvpacked_func)
        conv = spconv.SubMConv3d(32, 64, kernel_size=3, bias=True, indice_key="stage0")
        assert tuple(conv.weight.shape) == (64, 3, 3, 3, 32) and tuple(conv.bias.shape) == (64,)
        assert is_spconv_module(conv) and not is_spconv_module(torch.nn.Linear(2, 2))
        import inspect
        sig = inspect.signature(flash_attn.flash_attn_varlen_qkvpacked_func)
        assert list(sig.parameters)[:5] == ["qkv", "cu_seqlens", "max_seqlen", "dropout_p", "softmax_scale"]
        assert list(inspect.signature(torch_scatter.segment_csr).parameters) == ["src", "indptr", "out", "reduce"]
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.skipif(not os.path.isdir("/root/reference/pointcept"), reason="reference tree only exists in the build container")
def test_unmodified_reference_model_builds_on_the_stand_ins():
    """The reference's own PT-v3m1 file, imported unmodified with compat.install(force=True), builds its model out of
    this package's SubMConv3d (same state_dict as the reference layout and as our PT-v3m1).  Build container only."""
    import subprocess
    import sys
    code = r"""
import sys
sys.path.insert(0, %r)
from scenesplat_b200 import compat
compat.install(force=True)
from oracle import ref_shim
ref = ref_shim.load_reference()
import inspect
src = inspect.getsourcefile(ref.PointTransformerV3)
assert src.startswith("/root/reference/"), src
m = ref.PointTransformerV3(**ref_shim.LANG_BACKBONE_CFG)
assert type(m.embedding.stem.conv).__module__ == "scenesplat_b200.spconv_compat"
assert ref.ptv3.flash_attn.flash_attn_varlen_qkvpacked_func.__module__ == "scenesplat_b200.compat"
assert ref.ptv3.torch_scatter.segment_csr.__module__ == "scenesplat_b200.compat"
import scenesplat_b200 as S
ours = S.build_model(dict(type="PT-v3m1", **ref_shim.LANG_BACKBONE_CFG))
a = {k: tuple(v.shape) for k, v in m.state_dict().items()}
b = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
assert a == b, set(a) ^ set(b)
ours.load_state_dict(m.state_dict(), strict=True)
print("OK", len(a))
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


@pytest.mark.parametrize("case", ["mixed", "no_valid", "no_class"])
def test_training_losses_without_host_decisions_match_the_oracle(case):
    """scenesplat_b200/training.py evaluates the three lang losses with masks instead of boolean indexing / nonzero
    (nothing synchronises the stream between forward and backward).  Values AND gradients must equal the oracle's
    restatement of the reference (losses/misc.py:254-421), which removes rows / classes explicitly.  Pure torch: CPU."""
    from oracle import lang as olang
    from scenesplat_b200 import training as T
    g = torch.Generator().manual_seed(11)
    n, c, nc = 3000, 48, 16
    pred0 = torch.nn.functional.normalize(torch.randn(n, c, generator=g), dim=1)
    target = torch.nn.functional.normalize(torch.randn(n, c, generator=g), dim=1)
    mask = torch.rand(n, generator=g) < 0.7
    segment = torch.randint(-1, 9, (n,), generator=g)
    segment[segment == 7] = 8                     # class 7 absent, class 8 present
    segment[(segment == 3) & (torch.rand(n, generator=g) < 0.9)] = 2   # class 3 left with < 100 points
    if case == "no_valid":
        mask = torch.zeros(n, dtype=torch.bool)
    if case == "no_class":
        segment = torch.full((n,), -1)
    half = (torch.rand(n, generator=g) < 0.5).long()
    valid = mask & (segment != -1)

    def ours(p):
        sums, counts = T.class_half_sums(p.float(), valid, segment, half, nc)
        return (T.cosine_loss(p, target, mask, 1.0) + T.l2_loss(p, target, mask, 0.5)
                + 0.02 * T.contrastive_from_sums(sums, counts, nc, 0.2, "mean"))

    def ref(p):
        _, A, B = olang.class_half_sums(p, mask, segment, half)
        return (olang.cosine_loss(p, target, mask, 1.0) + olang.l2_loss(p, target, mask, 0.5)
                + olang.contrastive_from_sums(A, B, 0.2, 0.02))

    pa, pb = pred0.clone().requires_grad_(True), pred0.clone().requires_grad_(True)
    la, lb = ours(pa), ref(pb)
    np.testing.assert_allclose(la.item(), lb.item(), rtol=1e-5, atol=1e-6)
    la.backward()
    if lb.requires_grad:
        lb.backward()
        want = pb.grad
    else:
        want = torch.zeros_like(pred0)
    assert torch.isfinite(pa.grad).all()
    np.testing.assert_allclose(pa.grad.numpy(), want.numpy(), rtol=1e-4, atol=1e-7)


def test_cpe_conv_linear_folding_is_the_same_function():
    """ptv3.cpe_folded composes the xCPE conv with its Linear into per-tap weights (inference).  Against the oracle's
    fp32 conv followed by the Linear on the same inputs: identical up to the bf16 rounding of the composed weights
    (2^-9 relative per weight -> < 1e-2 relative L2 of the output).  Pure torch on the CPU."""
    from oracle import gridsample as ogs
    from oracle import subm_conv as oconv
    from scenesplat_b200 import ptv3, synthetic
    from scenesplat_b200.spconv_compat import SubMConv3d
    d = synthetic.chunk(3000, L=2.0, H=1.5, seed=3)
    g = ogs.grid_sample_train(d["coord"], 0.04)["grid_coord"]
    n, c = g.shape[0], 32
    torch.manual_seed(0)
    conv = SubMConv3d(c, c, kernel_size=3, bias=True)
    lin = torch.nn.Linear(c, c)
    x = torch.randn(n, c)
    nbr = oconv.kernel_map(g, np.zeros(n, dtype=np.int64), 3)
    with torch.no_grad():
        want = lin(oconv.subm_conv(x, nbr, conv.weight, conv.bias))
        w, b = ptv3.cpe_folded(conv, lin)                       # [k3, C, C] bf16, [C] fp32
        assert w.shape == (27, c, c) and w.dtype == torch.bfloat16 and b.dtype == torch.float32
        w5 = w.float().permute(1, 0, 2).reshape(c, 3, 3, 3, c)  # back to the conv layout [Cout, k, k, k, Cin]
        got = oconv.subm_conv(x, nbr, w5, b)
    rel = float((got - want).norm() / want.norm())
    assert rel < 1e-2, rel
    # the cache follows parameter updates
    with torch.no_grad():
        lin.weight.mul_(2.0)
    w2, _ = ptv3.cpe_folded(conv, lin)
    assert not torch.equal(w2, w)


def _write_scene_folder(path, n=5000, seed=0, with_segment=True):
    rng = np.random.default_rng(seed)
    os.makedirs(path, exist_ok=True)
    arrays = dict(coord=rng.random((n, 3)) * 5,                       # float64 on disk: get_data casts
                  color=rng.integers(0, 256, (n, 3)).astype(np.uint8), opacity=rng.random(n).astype(np.float32),
                  quat=rng.standard_normal((n, 4)).astype(np.float32), scale=(rng.random((n, 3)) * 3).astype(np.float32),
                  lang_feat=rng.standard_normal((n, 768)).astype(np.float16), valid_feat_mask=(rng.random(n) < 0.8).astype(np.uint8),
                  pc_coord=rng.random((n // 3, 3)).astype(np.float32), pc_segment20=rng.integers(-1, 20, n // 3))
    if with_segment:
        arrays["segment20"] = rng.integers(-1, 20, (n, 1)).astype(np.int64)
    for k, v in arrays.items():
        np.save(os.path.join(path, k + ".npy"), v)
    np.save(os.path.join(path, "unrelated.npy"), np.zeros(3))
    return arrays


@pytest.mark.parametrize("with_segment", [True, False])
def test_packed_scene_equals_reference_get_data(tmp_path, with_segment):
    """scene_io.pack_scene + load_scene returns exactly what the reference's get_data returns for the same folder
    (oracle/scene_io.py restates scannetgs.py:59-150): same keys, dtypes, shapes, bytes.  Sections are 4 KiB aligned,
    partial loads read only the requested attributes, a truncated file fails loudly."""
    from oracle import scene_io as oio
    from scenesplat_b200 import scene_io as sio
    folder = str(tmp_path / "scene0000_00")
    _write_scene_folder(folder, with_segment=with_segment)
    want = oio.get_data(folder, is_train=False)
    packed = str(tmp_path / "scene0000_00.sspk")
    size = sio.pack_scene(folder, packed)
    assert size == os.path.getsize(packed) and size % sio.ALIGN == 0
    header, meta = sio.read_header(packed)
    assert header % sio.ALIGN == 0 and all(e["offset"] % sio.ALIGN == 0 for e in meta["arrays"])
    got = sio.load_scene(packed, pinned=False).numpy()
    assert set(got) == set(want), (sorted(got), sorted(want))
    for k in want:
        assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape, k
        np.testing.assert_array_equal(got[k], want[k])
    assert got["scale"].max() <= 1.5 and got["lang_feat"].dtype == np.float16 and got["opacity"].shape[1] == 1
    part = sio.load_scene(packed, keys=("coord", "quat"), pinned=False)
    assert set(part.arrays) == {"coord", "quat"} and part.host.numel() < size // 4      # lang_feat was not read
    np.testing.assert_array_equal(part.numpy()["quat"], want["quat"])
    with pytest.raises(KeyError):
        sio.load_scene(packed, keys=("nope",), pinned=False)
    with open(packed, "r+b") as f:
        f.truncate(size - 8192)
    with pytest.raises(IOError):
        sio.load_scene(packed, pinned=False)


def test_ssl_variant_state_dict_matches_reference_golden_keys(golden):
    """PT-v3m1-simdino: identical parameter names / shapes / order as the reference's PointTransformerV3_SIMDINO
    (fixture generated from the unmodified reference file) and strict loading."""
    import scenesplat_b200 as S
    from tests.golden.make_golden import SMALL_CFG
    g = golden("ptv3_ssl.npz")
    ref_keys = [k[3:] for k in g.files if k.startswith("sd.")]
    model = S.build_model(dict(type="PT-v3m1-simdino", do_mask=True, pooling_reduce="max", **SMALL_CFG))
    sd = model.state_dict()
    assert list(sd.keys()) == ref_keys
    for k in ref_keys:
        assert tuple(sd[k].shape) == tuple(g["sd." + k].shape), k
    assert ref_keys[0] == "mask_token" and all(m.reduce == "max" for m in model.modules() if isinstance(m, S.SerializedPooling))
    assert not hasattr(S.PointTransformerV3SimDINO(**SMALL_CFG, do_mask=False), "dec")  # decoder only with do_mask (ref :676)
