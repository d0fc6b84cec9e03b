"""GPU gradient parity of the training path (scenesplat_b200/training.py) against torch autograd through the CPU
oracle (oracle/ptv3.py: ptv3_forward_autograd, fp32, train-mode BatchNorm).

Tolerance: the product runs bf16 operands in every GEMM / conv / attention, forward and backward (the reference
trains under fp16 AMP), the oracle is fp32 end to end; per-parameter gradients must agree in direction (cosine >
0.98) and size (relative L2 < 0.15), and the median relative L2 over all parameters must be < 0.05."""
import numpy as np
import pytest
import torch

from oracle import gridsample as ogs
from oracle import ptv3 as optv3
from scenesplat_b200 import synthetic

pytestmark = pytest.mark.gpu

TRAIN_CFG = dict(
    in_channels=11, order=("z", "z-trans", "hilbert", "hilbert-trans"), stride=(2, 2, 2),
    enc_depths=(1, 1, 1, 2), enc_channels=(32, 64, 64, 128), enc_num_head=(2, 4, 4, 8),
    enc_patch_size=(64, 64, 64, 64),
    dec_depths=(1, 1, 1), dec_channels=(96, 64, 64), dec_num_head=(2, 4, 4), dec_patch_size=(64, 64, 64),
    mlp_ratio=4, qkv_bias=True, drop_path=0.0, shuffle_orders=True, enable_flash=True,
    upcast_attention=False, upcast_softmax=False,
)


def _inputs(n_raw=5000, seed=5):
    d = synthetic.chunk(n_raw, L=1.6, H=1.2, seed=seed)
    res = ogs.grid_sample_train(d["coord"], 0.02)
    idx = res["idx_unique"]
    feat = synthetic.feat_from({k: v[idx] for k, v in d.items()})
    coord = d["coord"][idx]
    n = coord.shape[0]
    offset = np.array([n // 3, n], dtype=np.int64)
    return coord, res["grid_coord"], feat, offset


def test_backward_matches_oracle_autograd():
    import scenesplat_b200 as S
    coord, grid_coord, feat, offset = _inputs()
    n = coord.shape[0]
    torch.manual_seed(0)
    model = S.PointTransformerV3(**TRAIN_CFG)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    G = torch.randn(n, TRAIN_CFG["dec_channels"][0], generator=torch.Generator().manual_seed(3))

    # ---- gradient oracle (CPU, fp32)
    torch.manual_seed(7)
    perms = [torch.randperm(4).numpy() for _ in range(4)]
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in sd0.items()}
    out_ref = optv3.ptv3_forward_autograd(sd, TRAIN_CFG, coord, grid_coord, feat, offset, perms=perms)
    (out_ref * G).sum().backward()
    ref_grads = {k: v.grad for k, v in sd.items() if getattr(v, "requires_grad", False) and v.grad is not None}

    # ---- product (GPU, bf16 operands)
    model = model.cuda().train()
    data = dict(coord=torch.from_numpy(coord).cuda(), grid_coord=torch.from_numpy(grid_coord).cuda(),
                feat=torch.from_numpy(feat).cuda(), offset=torch.from_numpy(offset).cuda())
    torch.manual_seed(7)
    out = model(data).feat
    assert out.requires_grad and out.dtype == torch.float32
    fwd_rel = ((out.detach().cpu() - out_ref.detach()).norm() / out_ref.detach().norm()).item()
    assert fwd_rel < 3e-2, fwd_rel
    (out * G.cuda()).sum().backward()

    rels, bad = [], []
    gmax = max(float(v.norm()) for v in ref_grads.values())
    for name, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
        g, r = p.grad.detach().float().cpu().flatten(), ref_grads[name].flatten()
        if float(r.norm()) < 1e-4 * gmax:
            # structurally zero gradient: a bias whose only consumers are train-mode BatchNorms (a constant shift is
            # removed by the batch mean), e.g. down.proj.bias or the last fc2.bias of an encoder stage; both sides ~ 0
            assert float(g.norm()) < 2e-2 * gmax, (name, float(g.norm()), gmax)
            continue
        rel = ((g - r).norm() / r.norm().clamp(min=1e-12)).item()
        cos = torch.nn.functional.cosine_similarity(g, r, dim=0).item()
        rels.append(rel)
        if not (rel < 0.15 and cos > 0.98):
            bad.append((name, rel, cos))
    assert not bad, bad[:10]
    assert float(np.median(rels)) < 0.05, float(np.median(rels))
    # train-mode BatchNorm updated its running statistics (momentum 0.01)
    bn = model.embedding.stem.norm
    assert not torch.allclose(bn.running_mean.cpu(), sd0["embedding.stem.norm.running_mean"])
    assert int(bn.num_batches_tracked) == 1


def test_lang_pretrainer_training_step():
    """One optimisation step of LangPretrainer with the three lang losses: finite loss, finite gradients for every
    parameter, parameters move."""
    import scenesplat_b200 as S
    coord, grid_coord, feat, offset = _inputs(seed=9)
    n = coord.shape[0]
    cfg = dict(TRAIN_CFG, type="PT-v3m1", dec_channels=(768, 64, 64), dec_num_head=(16, 4, 4), drop_path=0.3)
    torch.manual_seed(0)
    model = S.LangPretrainer(backbone=cfg, criteria=[dict(type="CosineSimilarity"), dict(type="L2Loss"),
                                                      dict(type="AggregatedContrastiveLoss", schedule="all")]).cuda().train()
    rng = np.random.default_rng(0)
    lang = torch.nn.functional.normalize(torch.randn(n, 768), dim=1)
    data = dict(coord=torch.from_numpy(coord).cuda(), grid_coord=torch.from_numpy(grid_coord).cuda(),
                feat=torch.from_numpy(feat).cuda(), offset=torch.from_numpy(offset).cuda(),
                lang_feat=lang.cuda(), valid_feat_mask=torch.from_numpy(rng.random(n) < 0.8).cuda(),
                segment=torch.from_numpy(rng.integers(-1, 6, n)).cuda(), epoch_progress=0.9)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    before = {k: v.detach().clone() for k, v in model.named_parameters()}
    loss = model(data)["loss"]
    assert torch.isfinite(loss)
    loss.backward()
    for name, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
    opt.step()
    moved = sum(int(not torch.equal(before[k], v.detach())) for k, v in model.named_parameters())
    assert moved == len(before)


@pytest.mark.parametrize("c,xdt,dydt", [(32, torch.float32, torch.float32), (768, torch.float32, torch.bfloat16),
                                        (512, torch.bfloat16, torch.float32), (1024, torch.bfloat16, torch.bfloat16),
                                        (96, torch.float32, torch.float32)])
def test_layernorm_backward_kernel(c, xdt, dydt):
    """csrc/backward.cu vs torch autograd of F.layer_norm in fp32 on the same (possibly bf16-rounded) inputs:
    fp32 arithmetic on both sides, 2e-4 relative (1e-2 on dx when dx is stored as bf16)."""
    from scenesplat_b200 import ops
    torch.manual_seed(c)
    n = 3001
    x = (torch.randn(n, c) * 2 + 0.5).to(xdt)
    dy = torch.randn(n, c).to(dydt)
    g, b = torch.rand(c) + 0.5, torch.randn(c)
    xr = x.float().requires_grad_(True)
    gr = g.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (c,), gr, br, 1e-5).backward(dy.float())
    dx, dg, db = ops.layernorm_backward(x.cuda(), dy.cuda(), g.cuda(), 1e-5)
    assert dx.dtype == xdt
    tol = 1e-2 if xdt == torch.bfloat16 else 2e-4
    assert ((dx.float().cpu() - xr.grad).norm() / xr.grad.norm()).item() < tol
    assert ((dg.cpu() - gr.grad).norm() / gr.grad.norm()).item() < 2e-4
    assert ((db.cpu() - br.grad).norm() / br.grad.norm()).item() < 2e-4


@pytest.mark.parametrize("c", [32, 64, 256, 768])
def test_conv_backward_kernels(c):
    """SubMConvFn (own forward, own dgrad on the mirrored taps, own tensor-core wgrad) vs torch autograd through
    the fp32 oracle conv on the same bf16-rounded operands: bf16 products, fp32 accumulation -> 1e-2 relative L2."""
    from oracle import serialization as oser
    from oracle import subm_conv as oconv
    from scenesplat_b200 import ops, training
    d = synthetic.chunk(6000 if c > 256 else 12000, L=3.0, H=2.0, seed=2)
    res = ogs.grid_sample_train(d["coord"], 0.02)
    g = res["grid_coord"]
    n = g.shape[0]
    offset = np.array([n // 3, n], dtype=np.int64)
    batch = oser.offset2batch(offset)
    code, order, inv, depth = oser.serialization(g, batch, 2, ("z",))
    torch.manual_seed(1)
    x = torch.randn(n, c).bfloat16()
    w = (torch.randn(c, 3, 3, 3, c) * (1.0 / (c * 7) ** 0.5)).bfloat16().float()
    b = torch.randn(c)
    G = torch.randn(n, c).bfloat16()
    # oracle
    xr, wr, br = x.float().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    nbr_ref = oconv.kernel_map(g, batch, 3)
    (oconv.subm_conv(xr, nbr_ref, wr, br) * G.float()).sum().backward()
    # product
    dev = lambda a: torch.as_tensor(a).cuda()
    nbr, cnt = ops.kmap_build(dev(g), dev(batch), dev(code[0]), dev(order[0]), depth, 0, 3)
    pairs = ops.kmap_pairs(nbr, dev(order[0]), 3, cnt.cpu().numpy())
    xg, wg, bg = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    y = training.SubMConvFn.apply(xg, wg, bg, pairs, n)
    (y.float() * G.cuda().float()).sum().backward()
    for name, got, want in (("dx", xg.grad, xr.grad), ("dw", wg.grad, wr.grad), ("db", bg.grad, br.grad)):
        rel = ((got.float().cpu() - want).norm() / want.norm()).item()
        assert rel < 1e-2, (name, rel)


@pytest.mark.parametrize("H,d,K", [(2, 16, 256), (3, 32, 1024), (2, 48, 1024), (4, 48, 384)])
def test_patch_attention_backward_kernel(H, d, K):
    """tcgen05 attention backward (csrc/attention_bwd.cu) vs torch autograd through the fp32 oracle on the same bf16
    inputs: short item, ragged items (window rule on the last patch), full patches.  Also checks the log-sum-exp the
    training forward emits.  Tolerance: P, dS and the operands are bf16 (2^-9 relative each), outputs rounded to bf16:
    cosine > 0.999 and relative L2 < 2e-2 per gradient block (dq, dk, dv)."""
    from oracle import attention as oattn
    from scenesplat_b200 import ops
    rng = np.random.default_rng(3)
    offset = np.array([K // 2 + 3, K // 2 + 3 + 2 * K + 17, 4 * K + 40 + 333], dtype=np.int64)
    n, C = int(offset[-1]), H * d
    torch.manual_seed(0)
    qkv = (torch.randn(n, 3 * C) * 1.2).bfloat16()
    dout = torch.randn(n, C).bfloat16()
    order = np.concatenate([rng.permutation(np.arange(a, b)) for a, b in zip([0, *offset[:-1]], offset)])
    inverse = np.empty(n, dtype=np.int64)
    inverse[order] = np.arange(n)
    scale = d ** -0.5
    xr = qkv.float().requires_grad_(True)
    want = oattn.serialized_attention_core(xr, order, inverse, offset, K, H, scale)
    want.backward(dout.float())
    table = ops.patch_table(torch.from_numpy(offset).cuda(), K, n)
    ord_d = torch.from_numpy(order).cuda()
    out, lse2 = ops.patch_attention_lse(qkv.cuda(), ord_d, table, K, H, scale)
    # the forward output is unchanged by the extra lse store
    ref_out = ops.patch_attention(qkv.cuda(), ord_d, table, K, H, scale, impl="tc")
    assert torch.equal(out, ref_out)
    # log2-domain log-sum-exp by sorted position, against the oracle's scores
    pad, unpad, cu = oattn.patch_table(offset, K)
    q, k, _ = qkv.float().reshape(n, 3, H, d).unbind(1)
    srt = torch.from_numpy(order[pad])
    lse_ref = torch.empty(H, len(pad))
    for s0, e0 in zip(cu[:-1], cu[1:]):
        sc = torch.einsum("ihd,jhd->hij", q[srt[s0:e0]], k[srt[s0:e0]]) * scale
        lse_ref[:, s0:e0] = torch.logsumexp(sc, dim=-1) / np.log(2.0)
    lse_sorted = lse_ref[:, torch.from_numpy(unpad)]  # unpad: sorted position -> padded position of its own query row
    np.testing.assert_allclose(lse2.cpu().numpy(), lse_sorted.numpy(), atol=2e-2)  # weights are cut to bf16: 2^-8 relative
    dqkv = ops.patch_attention_backward(qkv.cuda(), out, dout.cuda(), lse2, ord_d, table, K, H, scale)
    torch.cuda.synchronize()
    got = dqkv.float().cpu()
    assert torch.isfinite(got).all()
    for name, sl in (("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))):
        a, b = got[:, sl].flatten(), xr.grad[:, sl].flatten()
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
        rel = float((a - b).norm() / b.norm())
        assert cos > 0.999 and rel < 2e-2, f"{name}: cos {cos:.5f} rel {rel:.4f}"


@pytest.mark.parametrize("n,c", [(1, 8), (5000, 32), (20001, 768), (3000, 3072), (777, 2304)])
def test_colsum_kernel_and_linear_fn(n, c):
    """csrc/backward.cu column sums (Linear bias gradient) vs a float64 sum of the same bf16 values (fp32 accumulation:
    1e-5 relative to the column's absolute sum), bit-identical between runs; LinearFn gradients vs torch autograd of
    F.linear on the same bf16 operands."""
    from scenesplat_b200 import ops
    from scenesplat_b200 import training as T
    torch.manual_seed(n + c)
    x = torch.randn(n, c).bfloat16()
    got = ops.colsum(x.cuda())
    again = ops.colsum(x.cuda())
    assert torch.equal(got, again)
    want = x.double().sum(0)
    tol = 1e-5 * x.double().abs().sum(0) + 1e-6
    assert bool(((got.cpu().double() - want).abs() <= tol).all())
    if c <= 768:
        lin = torch.nn.Linear(c, 64).cuda()
        xa = x.cuda().requires_grad_(True)
        xb = x.cuda().requires_grad_(True)
        g = torch.randn(n, 64, device="cuda").bfloat16()
        T._lin(lin, xa).backward(g)
        ga = [xa.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone()]
        lin.zero_grad()
        torch.nn.functional.linear(xb, lin.weight.to(torch.bfloat16), lin.bias.to(torch.bfloat16)).backward(g)
        gb = [xb.grad, lin.weight.grad, lin.bias.grad]
        for a, b, name in zip(ga, gb, ("dx", "dW", "db")):
            assert a.dtype == b.dtype and a.shape == b.shape, name
            # dx / dW: identical cuBLASLt GEMMs; db: fp32 column sum vs torch's bf16-rounded reduction
            np.testing.assert_allclose(a.float().cpu().numpy(), b.float().cpu().numpy(), rtol=1e-2, atol=1e-2 * float(b.float().abs().max()))


def test_stem_conv_wgrad_kernel():
    """csrc/conv_simt.cu stem weight gradient vs the gathered batched matmul in float64 (fp32 kernel: 1e-4 relative to
    the largest entry), bit-identical between runs (deterministic two-stage sum)."""
    from oracle import gridsample as ogs
    from oracle import serialization as oser
    from scenesplat_b200 import ops, synthetic
    d = synthetic.chunk(9000, L=3.0, H=2.0, seed=4)
    g = ogs.grid_sample_train(d["coord"], 0.02)["grid_coord"]
    n = g.shape[0]
    offset = np.array([n // 2, n], dtype=np.int64)
    batch = oser.offset2batch(offset)
    code, order, inv, depth = oser.serialization(g, batch, 2, oser.ORDERS)
    k, cin, cout = 5, 11, 32
    nbr, _ = ops.kmap_build(torch.from_numpy(g).cuda(), torch.from_numpy(batch).cuda(), torch.from_numpy(code[0]).cuda(),
                            torch.from_numpy(order[0]).cuda(), depth, 0, k)
    torch.manual_seed(0)
    x = torch.randn(n, cin)
    dy = torch.randn(n, cout)
    got = ops.stem_conv_wgrad(x.cuda(), dy.cuda(), nbr, k ** 3)
    assert torch.equal(got, ops.stem_conv_wgrad(x.cuda(), dy.cuda(), nbr, k ** 3))
    nb = nbr.cpu().long()
    xp = torch.cat([x.double(), torch.zeros(1, cin, dtype=torch.float64)], 0)
    idx = torch.where(nb >= 0, nb, torch.full_like(nb, n))
    want = torch.matmul(xp[idx].transpose(1, 2), dy.double())  # [k3, cin, cout]
    err = (got.cpu().double() - want).abs().max()
    assert err <= 1e-4 * want.abs().max(), float(err)


# ------------------------------------------------------------------------------------------------ adjoint kernels
def _pool_setup(n=5000, seed=3):
    """A pooling level on a random scene: (order0, seg_start, cluster, m) through the product's own index kernels."""
    from scenesplat_b200 import ops
    from oracle import serialization as oser
    rng = np.random.default_rng(seed)
    g = np.unique(rng.integers(0, 40, (n, 3)), axis=0)
    n = g.shape[0]
    batch = np.zeros(n, dtype=np.int64)
    code, order, inv, depth = oser.serialization(g, batch, 1, ("z", "hilbert"))
    dev = lambda a: torch.as_tensor(a).cuda()
    ix = ops.pool_index(dev(code), dev(order), dev(g), dev(batch), 1, [0, 1])
    return dev(order[0]).contiguous(), ix["seg_start"], ix["cluster"], ix["m"], n


@pytest.mark.parametrize("c", [32, 64, 256])
def test_segment_mean_and_unpool_adjoints(c):
    """SegmentMeanFn / UnpoolGatherAddFn (csrc/pool_loss_bwd.cu) against torch autograd through the same reductions
    written with index ops (fp32: equal up to summation order)."""
    from scenesplat_b200 import training
    order0, seg_start, cluster, m, n = _pool_setup()
    torch.manual_seed(0)
    src = torch.randn(n, c, device="cuda", requires_grad=True)
    w = torch.randn(m, c, device="cuda")
    out = training.SegmentMeanFn.apply(src, order0, seg_start, cluster)
    (out * w).sum().backward()
    src2 = src.detach().clone().requires_grad_(True)
    cnt = torch.bincount(cluster, minlength=m).float()
    ref = torch.zeros(m, c, device="cuda").index_add(0, cluster, src2) / cnt[:, None]
    (ref * w).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().cpu().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(src.grad.cpu().numpy(), src2.grad.cpu().numpy(), rtol=1e-6, atol=1e-7)

    a = torch.randn(n, c, device="cuda", requires_grad=True)
    child = torch.randn(m, c, device="cuda", requires_grad=True)
    wo = torch.randn(n, c, device="cuda")
    y = training.UnpoolGatherAddFn.apply(a, child, cluster, order0, seg_start)
    (y * wo).sum().backward()
    a2, child2 = a.detach().clone().requires_grad_(True), child.detach().clone().requires_grad_(True)
    y2 = a2 + child2[cluster]
    (y2 * wo).sum().backward()
    np.testing.assert_allclose(y.detach().cpu().numpy(), y2.detach().cpu().numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(a.grad.cpu().numpy(), a2.grad.cpu().numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(child.grad.cpu().numpy(), child2.grad.cpu().numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("tdtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("pdtype", [torch.float32, torch.bfloat16])
def test_cos_l2_loss_adjoint(pdtype, tdtype):
    """CosL2LossFn against torch autograd through the reference's formulation (losses/misc.py:254-295: boolean-mask
    gathers + nn.CosineSimilarity + squared distance), incl. fp16 targets (`lang_feat` as stored on disk) and the
    no-valid-row case."""
    from scenesplat_b200 import training
    torch.manual_seed(1)
    n, c = 3001, 768
    pred = torch.randn(n, c, device="cuda").to(pdtype).requires_grad_(True)
    target = torch.nn.functional.normalize(torch.randn(n, c, device="cuda"), dim=1).to(tdtype)
    mask = torch.rand(n, device="cuda") < 0.8
    up = torch.tensor(0.37, device="cuda")
    loss = training.cosine_loss(pred, target, mask, 1.3) + training.l2_loss(pred, target, mask, 0.7)
    (loss * up).backward()
    p2 = pred.detach().float().clone().requires_grad_(True)
    pv, tv = p2[mask], target.float()[mask]
    ref = 1.3 * (1 - torch.nn.functional.cosine_similarity(pv, tv, dim=1)).sum() / mask.sum() + \
        0.7 * ((pv - tv) ** 2).sum() / mask.sum()
    (ref * up).backward()
    assert abs(loss.item() - ref.item()) < 1e-4 * abs(ref.item()) + 1e-5
    tol = 1e-6 if pdtype == torch.float32 else 2e-3  # the bf16 gradient is the fp32 one rounded to bf16
    np.testing.assert_allclose(pred.grad.float().cpu().numpy(), p2.grad.cpu().numpy(), rtol=2e-2 if pdtype != torch.float32 else 1e-4,
                               atol=tol)
    # no valid row: loss 0, gradient 0
    p3 = pred.detach().clone().requires_grad_(True)
    l3 = training.cosine_loss(p3, target, torch.zeros_like(mask), 1.0)
    l3.backward()
    assert l3.item() == 0.0 and float(p3.grad.abs().max()) == 0.0


def test_class_half_sums_adjoint():
    from scenesplat_b200 import training
    torch.manual_seed(2)
    n, c, k = 4000, 768, 24
    pred = torch.randn(n, c, device="cuda", requires_grad=True)
    valid = torch.rand(n, device="cuda") < 0.7
    segment = torch.randint(-1, k + 3, (n,), device="cuda")  # incl. ignore (-1) and labels past the class table
    half = torch.randint(0, 2, (n,), device="cuda")
    w = torch.randn(2 * k, c, device="cuda")
    sums, counts = training.class_half_sums(pred, valid & (segment >= 0), segment, half, k)
    (sums * w).sum().backward()
    p2 = pred.detach().clone().requires_grad_(True)
    ok = valid & (segment >= 0) & (segment < k)
    key = (segment.clamp(0, k - 1) * 2 + half)
    ref = torch.zeros(2 * k, c, device="cuda").index_add(0, key, p2 * ok[:, None])
    (ref * w).sum().backward()
    np.testing.assert_allclose(sums.detach().cpu().numpy(), ref.detach().cpu().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_array_equal(counts.cpu().numpy(), torch.bincount(key[ok], minlength=2 * k).cpu().numpy())
    np.testing.assert_allclose(pred.grad.cpu().numpy(), p2.grad.cpu().numpy(), rtol=1e-6, atol=1e-6)
