"""Full-size (BASELINE.json configs[1]: one 299,277-voxel chunk) checks of the floating-point kernels through
size-independent properties, where the CPU oracle would take minutes: convexity and key-order invariance of the patch
attention, the softmax identities of its backward, and linearity / tap-sum identities of the tensor-core conv.
Tolerances are stated per check (bf16 operands / outputs: 2^-9 relative per rounding)."""
import numpy as np
import pytest
import torch

from scenesplat_b200 import synthetic

pytestmark = pytest.mark.gpu
N_FULL = 299277


def _attention_inputs(n, H, d, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    C = H * d
    qkv = (torch.randn(n, 3 * C, device="cuda", generator=g) * 1.2).bfloat16()
    order = torch.randperm(n, device="cuda", generator=g)
    return qkv, order


@pytest.mark.parametrize("H,d", [(16, 48), (2, 16)])
def test_attention_full_size_properties(H, d):
    from scenesplat_b200 import ops
    n, K, C = N_FULL, 1024, H * d
    qkv, order = _attention_inputs(n, H, d)
    table = ops.patch_table(torch.tensor([n], device="cuda"), K, n)
    scale = d ** -0.5
    out = ops.patch_attention(qkv, order, table, K, H, scale, impl="tc")
    assert torch.isfinite(out).all()
    # (1) every output row is a convex combination of its patch's V rows (the kernel normalises by the sum of the very
    #     weights it multiplies with): per channel within [min V, max V] of the patch, up to the bf16 rounding of the output
    P_full = n // K - 1                                   # full patches that lend no keys to the window patch
    srt = order[:P_full * K]
    v = qkv[srt, 2 * C:].float().view(P_full, K, C)
    o = out[srt].float().view(P_full, K, C)
    lo, hi = v.amin(1, keepdim=True), v.amax(1, keepdim=True)
    slack = 2.0 ** -8 * torch.maximum(lo.abs(), hi.abs()) + 1e-6
    assert bool(((o >= lo - slack) & (o <= hi + slack)).all())
    # (2) the order of the keys inside a patch is immaterial: shuffle every patch's rows among themselves (same patch
    #     membership, different gather order) -> same output per point up to summation order and the bf16 cut of P:
    #     |diff| <= 2e-2 and relative L2 < 5e-3
    perm = torch.argsort(torch.rand(n // K, K, device="cuda"), dim=1) + torch.arange(n // K, device="cuda")[:, None] * K
    order2 = order.clone()
    order2[:(n // K) * K] = order[perm.reshape(-1)]
    if n % K:                                             # keep the window patch's key set: leave its window untouched
        order2[n - K:] = order[n - K:]
        order2[(n // K - 1) * K:n - K] = order[(n // K - 1) * K:n - K]
    out2 = ops.patch_attention(qkv, order2, table, K, H, scale, impl="tc")
    diff = (out2.float() - out.float())
    assert float(diff.abs().max()) <= 2e-2 and float(diff.norm() / out.float().norm()) < 5e-3


def test_attention_backward_full_size_identities():
    """Softmax identities that hold for ANY inputs: sum_keys dV = sum_queries dO (rows of P sum to one) and
    <dQ, Q> = <dK, K> per head (rows of dS sum to zero).  fp32 accumulation over 299 k rows of bf16-rounded gradients:
    relative 1e-3 (dV) and 2e-3 (traces) of the absolute-value sums."""
    from scenesplat_b200 import ops
    n, K, H, d = N_FULL, 1024, 16, 48
    C = H * d
    qkv, order = _attention_inputs(n, H, d, seed=1)
    dout = torch.randn(n, C, device="cuda").bfloat16()
    table = ops.patch_table(torch.tensor([n], device="cuda"), K, n)
    out, lse2 = ops.patch_attention_lse(qkv, order, table, K, H, d ** -0.5)
    dqkv = ops.patch_attention_backward(qkv, out, dout, lse2, order, table, K, H, d ** -0.5)
    assert torch.isfinite(dqkv).all()
    dq, dk, dv = (dqkv[:, i * C:(i + 1) * C].double() for i in range(3))
    q, k = qkv[:, :C].double(), qkv[:, C:2 * C].double()
    sum_dv, sum_do = dv.sum(0), dout.double().sum(0)
    assert float((sum_dv - sum_do).abs().max()) <= 1e-3 * float(dout.double().abs().sum(0).max())
    tq = (dq * q).view(n, H, d).sum((0, 2))
    tk = (dk * k).view(n, H, d).sum((0, 2))
    scale_ref = (dq.abs() * q.abs()).view(n, H, d).sum((0, 2))
    assert bool(((tq - tk).abs() <= 2e-3 * scale_ref).all()), (tq, tk)


def test_conv_full_size_properties():
    """Tensor-core submanifold conv on the full chunk: with W_t = I for every tap the output is the sum of the active
    neighbours' rows (checked against torch gathers through the kernel map: bf16 products are exact, fp32 sums, bf16
    output: 2^-8 relative), and the conv is linear in its input (a second input doubles exactly in bf16)."""
    from scenesplat_b200 import GridSample, Point, ops
    from scenesplat_b200.spconv_compat import kernel_map_for
    d = synthetic.chunk(360000, seed=0)
    gs = GridSample(grid_size=0.02, hash_type="fnv", mode="train", keys=("coord",), return_grid_coord=True, device="cuda")
    s = gs({"coord": d["coord"]})
    g = torch.as_tensor(s["grid_coord"]).cuda()
    n, c = g.shape[0], 64
    assert n > 250000
    pt = Point(grid_coord=g, offset=torch.tensor([n], device="cuda"))
    pt.serialization(order=("z", "z-trans", "hilbert", "hilbert-trans"))
    ent = kernel_map_for(pt, 3, want_pairs=True)
    x = torch.randn(n, c, device="cuda").bfloat16()
    w = torch.eye(c, device="cuda").bfloat16().expand(27, c, c).contiguous()
    got = ops.subm_conv_gemm(x, ent["pairs"], w, None, n, out_dtype=torch.float32)
    nbr = ent["nbr"].long()                                # [27, n], -1 = no neighbour
    xp = torch.cat([x.float(), torch.zeros(1, c, device="cuda")], 0)
    want = torch.zeros(n, c, device="cuda")
    for t in range(27):
        want += xp[torch.where(nbr[t] >= 0, nbr[t], torch.full_like(nbr[t], n))]
    err = (got - want).abs()
    assert bool((err <= 2.0 ** -8 * want.abs() + 1e-5).all()), float(err.max())
    got2 = ops.subm_conv_gemm((x.float() * 2).bfloat16(), ent["pairs"], w, None, n, out_dtype=torch.float32)
    assert torch.equal(got2, got * 2)
    # every voxel is its own centre-tap neighbour and the kernel map is symmetric
    assert bool((nbr[13] == torch.arange(n, device="cuda")).all())
    cnt = (nbr >= 0).sum(0)
    assert int(cnt.min()) >= 1 and int((nbr >= 0).sum()) == ent["pairs"]["pairs"]
