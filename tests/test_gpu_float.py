"""GPU parity tests (through the C-ABI) of the floating-point kernels against the oracle.

Tolerances (stated per test): fp32 kernels 1e-4 relative; bf16 tensor-core kernels are compared with the
fp32 oracle evaluated on the SAME bf16-rounded inputs, so the remaining error is accumulation order +
one bf16 rounding of the result (2^-8 relative) + bf16 rounding of the per-tap products in the conv.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import attention as oattn
from oracle import gridsample as ogs
from oracle import lang as olang
from oracle import serialization as oser
from oracle import subm_conv as oconv
from scenesplat_b200 import synthetic

pytestmark = pytest.mark.gpu
ORDERS = oser.ORDERS


def dev(a):
    return torch.as_tensor(a).cuda()


def _scene(n_raw=20000, seed=2, nb=2, L=3.0):
    d = synthetic.chunk(n_raw, L=L, H=2.0, seed=seed)
    res = ogs.grid_sample_train(d["coord"], 0.02)
    g = res["grid_coord"]
    n = g.shape[0]
    offset = np.array([n // 3, n] if nb == 2 else [n], dtype=np.int64)
    batch = oser.offset2batch(offset)
    code, order, inv, depth = oser.serialization(g, batch, len(offset), ORDERS)
    return g, batch, offset, code, order, inv, depth


@pytest.mark.parametrize("k,cin,cout,bias", [(5, 11, 32, False), (3, 16, 16, True), (3, 32, 48, True)])
def test_subm_conv_simt(k, cin, cout, bias):
    from scenesplat_b200 import ops
    g, batch, offset, code, order, inv, depth = _scene(8000)
    torch.manual_seed(0)
    n = g.shape[0]
    x = torch.randn(n, cin)
    w = torch.randn(cout, k, k, k, cin) * 0.1
    b = torch.randn(cout) if bias else None
    nbr_ref = oconv.kernel_map(g, batch, k)
    want = oconv.subm_conv(x, nbr_ref, w, b)
    nbr, _ = ops.kmap_build(dev(g), dev(batch), dev(code[0]), dev(order[0]), depth, 0, k)
    wt = w.reshape(cout, k ** 3, cin).permute(1, 2, 0).contiguous().cuda()
    got = ops.subm_conv_simt(x.cuda(), nbr, wt, b.cuda() if bias else None)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-4)  # fp32 kernel
    # fused BN(eval)+GELU epilogue
    scale, shift = torch.rand(cout) + 0.5, torch.randn(cout)
    got2 = ops.subm_conv_simt(x.cuda(), nbr, wt, b.cuda() if bias else None, scale.cuda(), shift.cuda(), act=1)
    np.testing.assert_allclose(got2.cpu().numpy(), F.gelu(want * scale + shift).numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("c", [32, 64, 128, 256, 768])
def test_subm_conv_tensor_core(c, tile=256):
    """tcgen05 gather-GEMM conv (256-row tiles) against the fp32 oracle."""
    from scenesplat_b200 import ops
    g, batch, offset, code, order, inv, depth = _scene(6000 if c > 256 else 12000)
    torch.manual_seed(1)
    n = g.shape[0]
    x = (torch.randn(n, c)).bfloat16()
    w = (torch.randn(c, 3, 3, 3, c) * (1.0 / (c * 7) ** 0.5)).bfloat16()
    b = torch.randn(c)
    nbr_ref = oconv.kernel_map(g, batch, 3)
    want = oconv.subm_conv(x.float(), nbr_ref, w.float(), b)
    nbr, cnt = ops.kmap_build(dev(g), dev(batch), dev(code[0]), dev(order[0]), depth, 0, 3)
    pairs = ops.kmap_pairs(nbr, dev(order[0]), 3, cnt.cpu().numpy(), tile=tile)
    wk = w.reshape(c, 27, c).permute(1, 0, 2).contiguous().cuda()  # [27, cout, cin]
    got = ops.subm_conv_gemm(x.cuda(), pairs, wk, b.cuda(), n, out_dtype=torch.float32)
    err = (got.cpu() - want).abs().max().item()
    scale = want.abs().max().item()
    # bf16 products (2^-9 relative each, up to 27 summed) + fp32 accumulate: 2e-2 of the output scale
    assert err < 2e-2 * scale, (err, scale)
    rel = ((got.cpu() - want).norm() / want.norm()).item()
    assert rel < 5e-3, rel
    if tile == 256 and c >= 256:
        # the CTA-pair generation (csrc/conv_gemm3.cu, the default for C >= 256) and the single-CTA kernel compute the
        # same products in the same order: bit-identical
        got_other = ops.subm_conv_gemm(x.cuda(), pairs, wk, b.cuda(), n, out_dtype=torch.float32, impl="single")
        assert torch.equal(got_other, got)


@pytest.mark.parametrize("H,d,K,dtype", [(2, 16, 64, torch.float32), (4, 16, 1024, torch.bfloat16),
                                         (3, 32, 256, torch.bfloat16), (2, 48, 1024, torch.bfloat16),
                                         (1, 8, 32, torch.float32)])
def test_patch_attention_simt(H, d, K, dtype):
    from scenesplat_b200 import ops
    rng = np.random.default_rng(0)
    offset = np.array([K // 2 + 3, K // 2 + 3 + 2 * K + 17, 4 * K + 40], dtype=np.int64)  # short, ragged, exact-ish
    n = int(offset[-1])
    C = H * d
    torch.manual_seed(0)
    qkv = torch.randn(n, 3 * C).to(dtype)
    order = np.concatenate([rng.permutation(np.arange(a, b)) for a, b in zip([0, *offset[:-1]], offset)])
    inverse = np.empty(n, dtype=np.int64)
    inverse[order] = np.arange(n)
    scale = d ** -0.5
    want = oattn.serialized_attention_core(qkv.float(), order, inverse, offset, K, H, scale)
    table = ops.patch_table(dev(offset), K, n)
    got = ops.patch_attention(qkv.cuda(), dev(order), table, K, H, scale, out_dtype=torch.float32, impl="simt")
    tol = 1e-4 if dtype == torch.float32 else 2e-3  # bf16 inputs are exact here; fp32 math; exp2 vs exp
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=tol, atol=tol)


def test_add_layernorm_and_friends():
    from scenesplat_b200 import ops
    torch.manual_seed(0)
    for C in (16, 32, 96, 256, 768):
        n = 777
        res, delta = torch.randn(n, C), torch.randn(n, C)
        g0, b0, g1, b1 = torch.rand(C) + 0.5, torch.randn(C), torch.rand(C) + 0.5, torch.randn(C)
        y = res + F.layer_norm(delta, (C,), g0, b0, 1e-5)
        h = F.layer_norm(y, (C,), g1, b1, 1e-5)
        ro, no = ops.add_layernorm(res.cuda(), delta.cuda(), (g0.cuda(), b0.cuda()), (g1.cuda(), b1.cuda()),
                                   norm_dtype=torch.float32)
        np.testing.assert_allclose(ro.cpu().numpy(), y.numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(no.cpu().numpy(), h.numpy(), rtol=1e-4, atol=1e-4)
        # bf16 delta, no inner LN, bf16 normed output
        db = delta.bfloat16()
        ro, no = ops.add_layernorm(res.cuda(), db.cuda(), None, (g1.cuda(), b1.cuda()), norm_dtype=torch.bfloat16)
        y2 = res + db.float()
        np.testing.assert_allclose(ro.cpu().numpy(), y2.numpy(), rtol=1e-6, atol=1e-6)
        h2 = F.layer_norm(y2, (C,), g1, b1, 1e-5)
        np.testing.assert_allclose(no.float().cpu().numpy(), h2.numpy(), rtol=1e-2, atol=1e-2)  # bf16 output
    x = torch.randn(1000, 64)
    s, t = torch.rand(64) + 0.5, torch.randn(64)
    got = ops.affine_act(x.cuda(), s.cuda(), t.cuda(), act=1)
    # erf-GELU evaluated in float64 on the SAME fp32 affine result: the kernel's gelu_fast is within 3.3e-7 of it
    # (torch's own fp32 erf-GELU is only within 1.2e-6), so 1e-6 absolute
    want = F.gelu((x * s + t).double()).float()
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-6)
    xb = torch.randn(1000, 256).bfloat16()
    got = ops.affine_act(xb.cuda(), act=1)
    np.testing.assert_allclose(got.float().cpu().numpy(), F.gelu(xb.float()).bfloat16().float().numpy(), rtol=1e-2,
                               atol=1e-3)
    got = ops.l2_normalize(x.cuda())
    np.testing.assert_allclose(got.cpu().numpy(), F.normalize(x, dim=1).numpy(), rtol=1e-5, atol=1e-6)


def test_lang_head_and_losses_golden(golden):
    from scenesplat_b200 import ops
    g = golden("losses.npz")
    pred = torch.from_numpy(g["pred"].astype(np.float32)).cuda()
    target16 = torch.from_numpy(g["target"]).cuda()
    mask = torch.from_numpy(g["mask"]).cuda()
    seg = torch.from_numpy(g["segment"]).cuda()
    half = torch.from_numpy(g["half"]).cuda()
    text = torch.from_numpy(g["text"]).cuda()
    mx, lab = ops.lang_head_argmax(pred, text, impl="simt")
    np.testing.assert_allclose(mx.cpu().numpy(), g["max_prob"], rtol=1e-5, atol=1e-6)  # fp32
    np.testing.assert_array_equal(lab.cpu().numpy(), g["argmax"])
    # accumulate variant == dense sigmoid(logits)
    acc = torch.zeros(pred.shape[0], text.shape[0], device="cuda")
    ops.lang_head_accumulate(pred, text, acc, impl="simt")
    want = torch.sigmoid(pred @ text.t())
    np.testing.assert_allclose(acc.cpu().numpy(), want.cpu().numpy(), rtol=1e-5, atol=1e-6)
    # tensor-core head: bf16 operands (pred is fp16-exact but not bf16-exact; text rounds to bf16): logits move by
    # <= ~2^-9 * ||p|| * ||t|| = 2e-3, probs by a quarter of that -> 1e-3 absolute; labels may flip only on near ties
    mx_tc, lab_tc = ops.lang_head_argmax(pred, text, impl="tc")
    np.testing.assert_allclose(mx_tc.cpu().numpy(), g["max_prob"], atol=1e-3)
    assert (lab_tc.cpu().numpy() == g["argmax"]).mean() > 0.97
    acc_tc = torch.zeros(pred.shape[0], text.shape[0], device="cuda")
    ops.lang_head_accumulate(pred, text, acc_tc, impl="tc")
    np.testing.assert_allclose(acc_tc.cpu().numpy(), want.cpu().numpy(), atol=1e-3)
    for tgt in (target16, target16.float()):
        a = ops.cos_l2_sums(pred, tgt, mask).cpu().numpy()
        np.testing.assert_allclose(a[0] / a[2], g["cos"], rtol=1e-5)
        np.testing.assert_allclose(a[1] / a[2], g["l2"], rtol=1e-5)
        assert a[2] == g["mask"].sum()
    sums, counts = ops.class_half_sums(pred, mask, seg, half, 6)
    labs, A, B = olang.class_half_sums(pred.cpu(), mask.cpu(), seg.cpu(), half.cpu())
    for i, lab_ in enumerate(labs.tolist()):
        np.testing.assert_allclose(sums[2 * lab_].cpu().numpy(), A[i].numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(sums[2 * lab_ + 1].cpu().numpy(), B[i].numpy(), rtol=1e-4, atol=1e-5)
    con = olang.contrastive_from_sums(torch.stack([sums[2 * l] for l in labs.tolist()]).cpu(),
                                      torch.stack([sums[2 * l + 1] for l in labs.tolist()]).cpu(), 0.2, 0.025)
    np.testing.assert_allclose(con.numpy(), g["con"], rtol=1e-4)


@pytest.mark.parametrize("c", [128, 256, 512, 768])
def test_subm_conv_fused_add_layernorm(c):
    """conv gather-sum + LN + residual add + LN in one kernel (ss_subm_conv_reduce_add_ln) against the two-kernel path
    (ss_subm_conv_reduce -> fp32, then ss_add_layernorm): same products, same summation order; the only difference is
    the order of the LayerNorm reductions (warp shuffles vs. sub-group sums): 1e-5 relative."""
    from scenesplat_b200 import ops
    g, batch, offset, code, order, inv, depth = _scene(5000)
    torch.manual_seed(2)
    n = g.shape[0]
    x = torch.randn(n, c).bfloat16().cuda()
    w = (torch.randn(27, c, c) * (1.0 / (c * 7) ** 0.5)).bfloat16().cuda()
    b = torch.randn(c).cuda()
    res = (torch.randn(n, c) * 2).cuda()
    ln0 = (torch.rand(c).cuda() + 0.5, torch.randn(c).cuda())
    ln1 = (torch.rand(c).cuda() + 0.5, torch.randn(c).cuda())
    nbr, cnt = ops.kmap_build(dev(g), dev(batch), dev(code[0]), dev(order[0]), depth, 0, 3)
    pairs = ops.kmap_pairs(nbr, dev(order[0]), 3, cnt.cpu().numpy())
    z = ops.subm_conv_gemm(x, pairs, w, b, n, out_dtype=torch.float32)
    want_y, want_h = ops.add_layernorm(res, z, ln0, ln1, 1e-5, norm_dtype=torch.float32)
    got_y, got_h = ops.subm_conv_gemm_add_ln(x, pairs, w, b, res, ln0, ln1, 1e-5, inplace=False)
    np.testing.assert_allclose(got_y.cpu().numpy(), want_y.cpu().numpy(), rtol=2e-5, atol=2e-5)
    # bf16 output: equal up to one rounding step of the fp32 value it was rounded from
    np.testing.assert_allclose(got_h.float().cpu().numpy(), want_h.cpu().numpy(), rtol=2.0 ** -7, atol=1e-3)
    # in place
    r2 = res.clone()
    y2, h2 = ops.subm_conv_gemm_add_ln(x, pairs, w, b, r2, ln0, ln1, 1e-5, inplace=True)
    assert y2.data_ptr() == r2.data_ptr() and torch.equal(y2, got_y) and torch.equal(h2, got_h)


@pytest.mark.parametrize("c,n_raw,row", [(256, 5000, 0), (256, 70000, 2), (512, 40000, 1), (768, 70000, 3)])
def test_subm_conv_single_launch_equals_split(c, n_raw, row):
    """ss_subm_conv_fused_add_ln (gather-GEMM with reducer warps trailing it through per-tile counters, tiles taken in the
    order of their first output's rank) against ss_subm_conv_gemm_pair + ss_subm_conv_reduce_add_ln: the same products and
    the same summation order, so BIT-identical, on scenes with 1 .. 40 tiles per tap and along every serialized order."""
    from scenesplat_b200 import ops
    g, batch, offset, code, order, inv, depth = _scene(n_raw, seed=7 + row)
    torch.manual_seed(3)
    n = g.shape[0]
    x = torch.randn(n, c).bfloat16().cuda()
    w = (torch.randn(27, c, c) * (1.0 / (c * 7) ** 0.5)).bfloat16().cuda()
    b = torch.randn(c).cuda()
    res = (torch.randn(n, c) * 2).cuda()
    ln0 = (torch.rand(c).cuda() + 0.5, torch.randn(c).cuda())
    ln1 = (torch.rand(c).cuda() + 0.5, torch.randn(c).cuda())
    nbr, cnt = ops.kmap_build(dev(g), dev(batch), dev(code[row]), dev(order[row]), depth, row, 3)
    pairs = ops.kmap_pairs(nbr, dev(order[row]), 3, cnt.cpu().numpy())
    assert sorted(pairs["tile_order"].cpu().tolist()) == list(range(pairs["p_pad"] // 256))
    want_y, want_h = ops.subm_conv_gemm_add_ln(x, pairs, w, b, res, ln0, ln1, 1e-5, inplace=False, impl="split")
    for _ in range(3):  # the per-tile counters are reset by every call
        got_y, got_h = ops.subm_conv_gemm_add_ln(x, pairs, w, b, res, ln0, ln1, 1e-5, inplace=False, impl="fused")
        assert torch.equal(got_y, want_y) and torch.equal(got_h, want_h)


@pytest.mark.parametrize("logits", ["moderate", "huge", "hot_keys", "hot_rows"])
@pytest.mark.parametrize("H,d,K", [(2, 16, 1024), (3, 32, 1024), (2, 48, 1024), (4, 16, 256), (1, 48, 100)])
def test_patch_attention_tensor_core(H, d, K, logits):
    """tcgen05 kernel vs the fp32 oracle AND vs the independent SIMT kernel on the same bf16 inputs.
    `logits`: "moderate" = unit-scale inputs; "huge" = logits of several hundred (the running reference moves on most
    steps: lazy rescaling of O), "hot_keys" = a band of 100 sorted key positions per patch has 6x the norm (the
    reference jumps in the middle of a tile), "hot_rows" = every 5th row has 6x the norm (rows of one warp rescale at
    different steps).
    Tolerance (bf16 has 8 significant bits): the softmax weights are cut to bf16 before P.V (<= 2^-8 relative
    per weight, normalised by the sum of the SAME cut weights) and the result is rounded to bf16 (2^-9 relative):
    |err| <= 8e-3 + 2^-7 |want| per element (outputs reach |x| ~ 6 on peaked rows, where one bf16 ulp is 0.03),
    and 1e-2 in relative L2 over the tensor."""
    from scenesplat_b200 import ops
    rng = np.random.default_rng(1)
    offset = np.array([K // 2 + 3, K // 2 + 3 + 2 * K + 17, 4 * K + 40 + 333], dtype=np.int64)
    n = int(offset[-1])
    C = H * d
    torch.manual_seed(0)
    qkv = torch.randn(n, 3 * C) * (6.0 if logits == "huge" else 1.5)
    order = np.concatenate([rng.permutation(np.arange(a, b)) for a, b in zip([0, *offset[:-1]], offset)])
    inverse = np.empty(n, dtype=np.int64)
    inverse[order] = np.arange(n)
    if logits == "hot_keys":
        pos = np.arange(n)
        hot = order[(pos % K >= K // 3) & (pos % K < K // 3 + 100)]
        qkv[hot, C:2 * C] *= 6.0
    elif logits == "hot_rows":
        qkv[::5, :2 * C] *= 6.0
    qkv = qkv.bfloat16()
    scale = d ** -0.5
    want = oattn.serialized_attention_core(qkv.float(), order, inverse, offset, K, H, scale)
    table = ops.patch_table(dev(offset), K, n)
    got = ops.patch_attention(qkv.cuda(), dev(order), table, K, H, scale, impl="tc")
    simt = ops.patch_attention(qkv.cuda(), dev(order), table, K, H, scale, impl="simt")
    torch.cuda.synchronize()
    # the absolute term scales with |V| (weights cut to bf16 move the result by <= 2^-8 of the V rows they mix)
    # (peaked rows -- huge logits, hot keys -- put whole-ulp weight errors on single V rows: 1.2e-2 there)
    tol = {"moderate": 8e-3, "huge": 4e-2}.get(logits, 1.2e-2) + 2.0 ** -7 * want.abs()
    excess = ((got.float().cpu() - want).abs() - tol).max().item()
    assert excess <= 0, excess
    excess_simt = ((got.float() - simt.float()).abs().cpu() - tol).max().item()
    assert excess_simt <= 0, excess_simt
    rel = ((got.float().cpu() - want).norm() / want.norm()).item()
    assert rel < 1e-2, rel


@pytest.mark.timeout(120)
@pytest.mark.parametrize("n,nb,H,d", [(60000, 13, 16, 32), (60000, 13, 16, 48), (30000, 7, 16, 16), (20000, 31, 8, 48),
                                      (45000, 9, 12, 32), (3000, 3, 2, 16)])
def test_patch_attention_item_sequences(n, nb, H, d):
    """The persistent tcgen05 kernel walks many (head, patch) items per CTA; batches of `nb` elements put ragged patches
    (odd and even numbers of query tiles, short key windows, the last-patch window rule) BETWEEN full ones in every CTA's
    item sequence.  13 elements x (4 full patches + 5 query tiles) at d = 32 deadlocked the first persistent version (the
    groups' ping-pong barrier, see csrc/attention_tc.cu); the mbarrier waits trap after 20 s, so a protocol bug fails
    this test instead of hanging it.  Checked against the independent SIMT kernel on the same bf16 inputs (same tolerance
    as test_patch_attention_tensor_core) and, for the small case, against the fp32 oracle."""
    from scenesplat_b200 import ops
    K = 1024
    rng = np.random.default_rng(n + nb)
    offset = np.array([(n * (b + 1)) // nb for b in range(nb)], dtype=np.int64)
    C = H * d
    torch.manual_seed(1)
    qkv = (torch.randn(n, 3 * C) * 1.5).bfloat16()
    order = np.concatenate([rng.permutation(np.arange(a, b)) for a, b in zip([0, *offset[:-1]], offset)])
    scale = d ** -0.5
    table = ops.patch_table(dev(offset), K, n)
    got = ops.patch_attention(qkv.cuda(), dev(order), table, K, H, scale, impl="tc")
    simt = ops.patch_attention(qkv.cuda(), dev(order), table, K, H, scale, impl="simt")
    torch.cuda.synchronize()
    tol = 8e-3 + 2.0 ** -7 * simt.float().abs()
    excess = ((got.float() - simt.float()).abs() - tol).max().item()
    assert excess <= 0, excess
    if n <= 5000:
        inverse = np.empty(n, dtype=np.int64)
        inverse[order] = np.arange(n)
        want = oattn.serialized_attention_core(qkv.float(), order, inverse, offset, K, H, scale)
        assert ((got.float().cpu() - want).norm() / want.norm()).item() < 1e-2


@pytest.mark.parametrize("n,cin,cout", [(1, 16, 32), (300, 128, 32), (257, 256, 64), (5001, 3072, 768), (40001, 2048, 512)])
@pytest.mark.parametrize("inplace", [True, False])
def test_linear_residual_cta_pair(n, cin, cout, inplace):
    """fc2 + bias + residual add in the GEMM epilogue (csrc/gemm2cta.cu, RES = 1) vs float64 on the same bf16 operands:
    the fp32 result carries only the accumulation-order error of an fp32 dot product of `cin` bf16 products
    (|err| <= 2e-5 * sqrt(cin) * scale), the bf16 copy is its correctly rounded value."""
    from scenesplat_b200 import ops
    torch.manual_seed(n + cin)
    x = torch.randn(n, cin).bfloat16()
    w = (torch.randn(cout, cin) / cin ** 0.5).bfloat16()
    b = torch.randn(cout)
    res = torch.randn(n, cout) * 3
    want = res.double() + x.double() @ w.double().t() + b.double()
    r = res.cuda()
    out, shadow = ops.linear_residual(x.cuda(), w.cuda(), b.cuda(), r, inplace=inplace)
    assert (out.data_ptr() == r.data_ptr()) == inplace
    if not inplace:
        assert torch.equal(r.cpu(), res)  # untouched
    err = (out.double().cpu() - want).abs().max().item()
    assert err <= 2e-5 * cin ** 0.5 * 4, err
    assert torch.equal(shadow, out.bfloat16())
    out2, none = ops.linear_residual(x.cuda(), w.cuda(), b.cuda(), res.cuda(), want_bf16=False)
    assert none is None and torch.equal(out2, out)


@pytest.mark.parametrize("n,cin,cout", [(1, 16, 32), (255, 48, 96), (257, 64, 256), (1000, 768, 3072), (5001, 3072, 768),
                                        (70001, 512, 2048),
                                        # cin = 32: the row-streaming warp-MMA kernel (csrc/gemm_narrow.cu), incl. a strip
                                        # tail, one / two 128-column passes and a partial second pass
                                        (1, 32, 32), (255, 32, 96), (70001, 32, 128), (4099, 32, 256), (300, 32, 160)])
@pytest.mark.parametrize("act", [0, 1])
def test_linear_act_cta_pair(n, cin, cout, act):
    """csrc/gemm2cta.cu (tcgen05 cta_group::2, fused bias + exact GELU; cin = 32: csrc/gemm_narrow.cu) vs float64 on the
    same bf16 operands, incl. row / column / K tails.  bf16 output of an fp32 accumulation: |err| <= 2^-8 |want| + 1e-3 * sqrt(cin) * 2^-8 (accumulation
    order), and relative L2 < 4e-3."""
    from scenesplat_b200 import ops
    torch.manual_seed(n + cin)
    x = torch.randn(n, cin).bfloat16()
    w = (torch.randn(cout, cin) / cin ** 0.5).bfloat16()
    b = torch.randn(cout)
    got = ops.linear_act(x.cuda(), w.cuda(), b.cuda(), act).float().cpu()
    want = x.double() @ w.double().t() + b.double()
    if act:
        want = F.gelu(want)
    err = (got.double() - want).abs()
    assert bool((err <= 2.0 ** -8 * want.abs() + 2e-3).all()), float(err.max())
    assert float(err.norm() / want.norm()) < 4e-3
    nob = ops.linear_act(x.cuda(), w.cuda(), None, 0).float().cpu()          # bias is optional
    assert float((nob.double() - x.double() @ w.double().t()).abs().max()) <= 2.0 ** -7 * float(want.abs().max()) + 2e-3


@pytest.mark.parametrize("n,c,bf", [(1, 8, True), (1000, 768, True), (777, 256, False), (33, 40, True)])
def test_add_l2_normalize(n, c, bf):
    """ss_add_l2_normalize = F.normalize(res + delta, p=2, dim=1, eps=1e-12) in one pass (fp32): 2e-6 relative (the sum
    of squares is a warp-shuffle tree instead of torch's order; one reciprocal instead of a division)."""
    from scenesplat_b200 import ops
    torch.manual_seed(n + c)
    res = torch.randn(n, c).cuda() * 3
    delta = torch.randn(n, c).cuda()
    if bf:
        delta = delta.bfloat16()
    want = F.normalize(res + delta.float(), p=2, dim=1, eps=1e-12)
    got = ops.add_l2_normalize(res, delta, eps=1e-12)
    np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=2e-6, atol=1e-7)
    # the bf16 copy written by the same kernel is the rounding torch would produce, is what the head's GEMM operand
    # lookup returns, and is dropped once the fp32 tensor is modified in place
    sh = ops.bf16_shadow(got)
    assert sh is not None and sh.dtype == torch.bfloat16 and torch.equal(sh, got.bfloat16())
    assert ops._head_operand(got, False) is sh
    assert ops.bf16_shadow(ops.add_l2_normalize(res, delta, eps=1e-12, want_bf16=False)) is None
    got.mul_(2.0)
    assert ops.bf16_shadow(got) is None
    zero = ops.add_l2_normalize(torch.zeros(4, c).cuda(), torch.zeros(4, c).cuda().bfloat16(), eps=1e-12)
    assert float(zero.abs().max()) == 0.0  # eps guards the zero row, as in F.normalize
