"""GPU parity tests of the third-party stand-ins (scenesplat_b200/compat.py), called with the EXACT argument forms of
the reference's call sites (point_transformer_v3m1_base.py:189-196, :277-284, :416-421; structure.py:131-138).

Tolerances: segment_csr is fp32 arithmetic (1e-5); attention and the tensor-core conv are bf16 kernels, compared with
the fp32 oracle on the same bf16-rounded inputs: |err| <= 8e-3 + 2^-7 |want| per element (one bf16 rounding of the
output plus bf16 probabilities / per-tap products).  Gradients: cosine > 0.99 and relative L2 < 0.05 against torch
autograd through the fp32 oracle.
"""
import numpy as np
import pytest
import torch

from oracle import attention as oattn
from oracle import gridsample as ogs
from oracle import pooling as opool
from oracle import subm_conv as oconv
from scenesplat_b200 import synthetic

pytestmark = pytest.mark.gpu


def _close(got, want, what):
    got, want = got.float().cpu(), want.float().cpu()
    err = (got - want).abs()
    tol = 8e-3 + 2.0 ** -7 * want.abs()
    assert bool((err <= tol).all()), f"{what}: max err {err.max():.4g}"


def _grad_close(got, want, what):
    got, want = got.float().cpu().flatten(), want.float().cpu().flatten()
    cos = torch.dot(got, want) / (got.norm() * want.norm() + 1e-30)
    rel = (got - want).norm() / (want.norm() + 1e-30)
    assert cos > 0.99 and rel < 0.05, f"{what}: cos {cos:.4f} rel {rel:.4f}"


@pytest.mark.parametrize("reduce", ["sum", "mean", "max", "min"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_segment_csr_signature(reduce, dtype):
    from scenesplat_b200.compat import segment_csr
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 9, size=300)          # includes empty segments
    counts[0], counts[-1] = 0, 0
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    n = int(indptr[-1])
    src = torch.randn(n, 70).to(dtype)
    got = segment_csr(src.cuda(), torch.from_numpy(indptr).cuda(), reduce=reduce)
    assert got.dtype == dtype and got.shape == (300, 70)
    with np.errstate(invalid="ignore", divide="ignore"):
        want = opool.segment_csr(src.float().numpy(), np.arange(n), indptr, reduce)
    want = np.where(counts[:, None] > 0, want, 0.0)  # torch_scatter: empty segments give 0
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    np.testing.assert_allclose(got.float().cpu().numpy(), want, rtol=tol, atol=tol)


def test_segment_csr_reference_call_and_grad():
    """segment_csr(proj(feat)[indices], idx_ptr, reduce="mean") as in SerializedPooling.forward (:416-421)."""
    from scenesplat_b200.compat import segment_csr
    torch.manual_seed(1)
    n, c = 5000, 64
    cluster = torch.sort(torch.randint(0, 1500, (n,))).values
    _, cluster, counts = torch.unique(cluster, return_inverse=True, return_counts=True)
    perm = torch.randperm(n)
    cluster = cluster[perm]
    indices = torch.sort(cluster).indices
    idx_ptr = torch.cat([counts.new_zeros(1), torch.cumsum(counts, 0)])
    feat = torch.randn(n, c)
    x = feat.cuda().requires_grad_(True)
    got = segment_csr(x[indices.cuda()], idx_ptr.cuda(), reduce="mean")
    g = torch.randn(got.shape)
    got.backward(g.cuda())
    xr = feat.clone().requires_grad_(True)
    seg = torch.repeat_interleave(torch.arange(counts.numel()), counts)
    want = torch.zeros(counts.numel(), c).index_add_(0, seg, xr[indices]) / counts[:, None]
    want.backward(g)
    np.testing.assert_allclose(want.detach().numpy(), opool.segment_csr(feat.numpy(), indices.numpy(), idx_ptr.numpy(), "mean"),
                               rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), xr.grad.numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("H,d", [(4, 32), (2, 16), (3, 48)])
def test_flash_attn_varlen_signature(H, d):
    """flash_attn_varlen_qkvpacked_func(qkv.half().reshape(-1, 3, H, d), cu_seqlens, max_seqlen=K, dropout_p, softmax_scale)."""
    from scenesplat_b200.compat import flash_attn_varlen_qkvpacked_func
    torch.manual_seed(2)
    K = 1024
    lens = [1024, 1024, 300, 1024, 77, 1]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    T, C = int(cu[-1]), H * d
    qkv = (torch.randn(T, 3 * C) * 1.5).half()
    scale = d ** -0.5
    got = flash_attn_varlen_qkvpacked_func(qkv.cuda().reshape(-1, 3, H, d), torch.from_numpy(cu).cuda(), max_seqlen=K,
                                           dropout_p=0.0, softmax_scale=scale)
    assert got.shape == (T, H, d) and got.dtype == torch.float16
    xin = qkv.to(torch.bfloat16).float()          # the kernel computes on bf16 operands
    want = oattn.varlen_attention(xin, cu, H, scale)
    _close(got.reshape(T, C), want.reshape(T, C), "flash_attn stand-in")


def test_flash_attn_varlen_grad():
    from scenesplat_b200.compat import flash_attn_varlen_qkvpacked_func
    torch.manual_seed(3)
    H, d, K = 4, 32, 256
    lens = [256, 256, 100, 256, 31]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    T, C = int(cu[-1]), H * d
    qkv = (torch.randn(T, 3 * C)).to(torch.bfloat16)
    g = torch.randn(T, H, d)
    x = qkv.cuda().requires_grad_(True)
    out = flash_attn_varlen_qkvpacked_func(x.reshape(-1, 3, H, d), torch.from_numpy(cu).cuda(), K, 0.0, d ** -0.5)
    out.backward(g.cuda().to(out.dtype))
    xr = qkv.float().requires_grad_(True)
    want = oattn.varlen_attention(xr, cu, H, d ** -0.5)
    want.reshape(T, H, d).backward(g)
    _grad_close(x.grad, xr.grad, "d qkv")


def _voxels(n_raw=9000, seed=5):
    d = synthetic.chunk(n_raw, L=3.0, H=2.0, seed=seed)
    g = ogs.grid_sample_train(d["coord"], 0.02)["grid_coord"]
    n = g.shape[0]
    batch = np.zeros(n, dtype=np.int64)
    batch[n // 3:] = 1
    return g, batch


@pytest.mark.parametrize("k,cin,cout", [(3, 32, 64), (5, 11, 32)])
def test_spconv_signature_standalone_tensor(k, cin, cout):
    """SparseConvTensor(features, indices=[batch, grid_coord].int(), spatial_shape, batch_size) -> SubMConv3d(...)(x):
    a tensor built the spconv way, with no Point behind it."""
    from scenesplat_b200.compat import install
    mods = install(force=True)
    spconv = mods["spconv.pytorch"]
    assert mods["spconv.pytorch.modules"].is_spconv_module is not None
    g, batch = _voxels()
    n = g.shape[0]
    torch.manual_seed(4)
    feat = torch.randn(n, cin)
    indices = torch.cat([torch.from_numpy(batch)[:, None].int(), torch.from_numpy(g).int()], dim=1).contiguous()
    shape = (torch.from_numpy(g).max(0).values + 96).tolist()
    x = spconv.SparseConvTensor(features=feat.cuda(), indices=indices.cuda(), spatial_shape=shape, batch_size=2)
    conv = spconv.SubMConv3d(cin, cout, kernel_size=k, padding=1, bias=True, indice_key="stem").cuda()
    assert mods["spconv.pytorch.modules"].is_spconv_module(conv)
    with torch.no_grad():
        y = conv(x)
    assert y.indices is x.indices and y.features.shape == (n, cout)
    nbr = oconv.kernel_map(g, batch, k)
    w, b = conv.weight.detach().cpu(), conv.bias.detach().cpu()
    if cin % 16 == 0:      # tensor-core path: bf16 operands
        want = oconv.subm_conv(feat.to(torch.bfloat16).float(), nbr, w.to(torch.bfloat16).float(), b)
        _close(y.features, want, "SubMConv3d (tcgen05)")
    else:                  # fp32 SIMT path (the stem)
        want = oconv.subm_conv(feat, nbr, w, b)
        np.testing.assert_allclose(y.features.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-4)
    # a second conv on the same lineage reuses the kernel map (spconv: indice_key)
    y2 = conv(y.replace_feature(y.features[:, :cin].contiguous() if cout >= cin else feat.cuda()))
    assert y2._point is y._point


def test_spconv_signature_grad():
    from scenesplat_b200.compat import install
    spconv = install(force=True)["spconv.pytorch"]
    g, batch = _voxels(6000, seed=6)
    n, cin, cout = g.shape[0], 32, 32
    torch.manual_seed(5)
    feat = torch.randn(n, cin).to(torch.bfloat16).float()
    indices = torch.cat([torch.from_numpy(batch)[:, None].int(), torch.from_numpy(g).int()], dim=1).contiguous()
    conv = spconv.SubMConv3d(cin, cout, kernel_size=3, bias=True, indice_key="s").cuda()
    x = feat.cuda().requires_grad_(True)
    y = conv(spconv.SparseConvTensor(x, indices.cuda(), [500, 500, 500], 2)).features
    gy = torch.randn(n, cout).to(torch.bfloat16).float()
    y.backward(gy.cuda())
    w = conv.weight.detach().cpu().to(torch.bfloat16).float().requires_grad_(True)
    b = conv.bias.detach().cpu().clone().requires_grad_(True)
    xr = feat.clone().requires_grad_(True)
    want = oconv.subm_conv(xr, oconv.kernel_map(g, batch, 3), w, b)
    want.backward(gy)
    _close(y.detach(), want.detach(), "SubMConv3d forward under autograd")
    _grad_close(x.grad, xr.grad, "dx")
    _grad_close(conv.weight.grad, w.grad, "dW")
    _grad_close(conv.bias.grad, b.grad, "db")
