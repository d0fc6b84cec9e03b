"""Stand-ins with the EXACT call signatures of the three third-party packages the reference's PT-v3m1 imports
(SURVEY.md section 8b), so the unmodified reference file can run on this package's kernels:

    import scenesplat_b200.compat as compat
    compat.install()          # before `import pointcept`: fills sys.modules["spconv.pytorch"], ["torch_scatter"], ["flash_attn"]

  torch_scatter.segment_csr(src, indptr, out=None, reduce="sum")
        call site pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py:416-421 (reduce="mean")
  flash_attn.flash_attn_varlen_qkvpacked_func(qkv, cu_seqlens, max_seqlen, dropout_p, softmax_scale, ...)
        call site :189-196 (qkv.half().reshape(-1, 3, H, d), cu_seqlens int32, max_seqlen = patch_size)
  spconv.pytorch.{SubMConv3d, SparseConvTensor}, spconv.pytorch.modules.is_spconv_module
        call sites :277-284, :499-506, pointcept/models/utils/structure.py:131-138, pointcept/models/modules.py:64-75

All three work on CUDA tensors only (there is no CPU path) and carry gradients (scenesplat_b200/training.py).
"""
from __future__ import annotations

import sys
import types

import torch

from . import ops
from . import spconv_compat

_REDUCES = ("sum", "mean", "max", "min")


# ------------------------------------------------------------------------------------------- torch_scatter
class _SegmentCsrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, indptr, reduce):
        out = ops.segment_reduce(src, None, indptr, reduce, out_dtype=src.dtype)
        ctx.save_for_backward(indptr)
        ctx.n, ctx.reduce = src.shape[0], reduce
        return out

    @staticmethod
    def backward(ctx, dout):
        (indptr,) = ctx.saved_tensors
        cnt = indptr[1:] - indptr[:-1]
        seg = torch.repeat_interleave(torch.arange(cnt.numel(), device=dout.device), cnt, output_size=int(ctx.n))
        if ctx.reduce == "mean":
            dout = dout / cnt.clamp(min=1)[:, None].to(dout.dtype)
        return dout[seg], None, None


def segment_csr(src, indptr, out=None, reduce="sum"):
    """torch_scatter.segment_csr for the layout the reference uses: src [N, C] (fp32 / bf16 / fp16), indptr [M + 1]
    int64 along dim 0.  Empty segments give 0, as in torch_scatter."""
    if reduce not in _REDUCES:
        raise ValueError(f"segment_csr: unknown reduce {reduce!r}")
    if indptr.dim() != 1 or src.dim() < 1:
        raise NotImplementedError("segment_csr stand-in: indptr must be 1-D (segments along dim 0)")
    shape = src.shape
    x = src.reshape(shape[0], -1)
    back = None
    if x.dtype not in (torch.float32, torch.bfloat16):
        back, x = x.dtype, x.float()
    ip = indptr.to(torch.int64).contiguous()
    if torch.is_grad_enabled() and x.requires_grad:
        if reduce not in ("sum", "mean"):
            raise NotImplementedError("segment_csr stand-in: gradients for sum / mean only")
        res = _SegmentCsrFn.apply(x.contiguous(), ip, reduce)
    else:
        res = ops.segment_reduce(x, None, ip, reduce, out_dtype=x.dtype)
    if back is not None:
        res = res.to(back)
    res = res.reshape((ip.numel() - 1,) + tuple(shape[1:]))
    if out is not None:
        out.copy_(res)
        return out
    return res


# ------------------------------------------------------------------------------------------- flash_attn
def _plan_from_cu_seqlens(cu, K):
    """(start, n_full, tail_rows, end) runs for training._attention_backward: consecutive length-K sequences
    are one batched run, any shorter sequence attends to itself."""
    plan, i, n_seq = [], 0, len(cu) - 1
    while i < n_seq:
        s0, ln = cu[i], cu[i + 1] - cu[i]
        if ln == K:
            j = i
            while j < n_seq and cu[j + 1] - cu[j] == K:
                j += 1
            plan.append((s0, j - i, 0, cu[j]))
            i = j
        else:
            if ln > 0:
                plan.append((s0, 0, ln, cu[i + 1]))
            i += 1
    return plan


class _VarlenAttnFn(torch.autograd.Function):
    """tcgen05 forward (with log-sum-exp) and tcgen05 backward; shapes outside the kernels' range (head_dim not in
    {16, 32, 48} or max_seqlen > 1024) recompute with torch's fused attention."""

    @staticmethod
    def forward(ctx, qkv2d, table, cu_seqlens, K, H, scale):
        from .training import _own_attention_backward
        ident = _arange(qkv2d.shape[0], qkv2d.device)
        ctx.own = _own_attention_backward(qkv2d.shape[1] // (3 * H), K)
        if ctx.own:
            out, lse2 = ops.patch_attention_lse(qkv2d, ident, table, K, H, scale)
            ctx.save_for_backward(qkv2d, cu_seqlens, table, ident, out, lse2)
        else:
            out = ops.patch_attention(qkv2d, ident, table, K, H, scale)
            ctx.save_for_backward(qkv2d, cu_seqlens)
        ctx.K, ctx.H, ctx.scale = K, H, scale
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.own:
            qkv2d, cu, table, ident, out, lse2 = ctx.saved_tensors
            dqkv = ops.patch_attention_backward(qkv2d, out, dout.to(qkv2d.dtype), lse2, ident, table, ctx.K, ctx.H, ctx.scale)
        else:
            from .training import _attention_backward
            qkv2d, cu = ctx.saved_tensors
            plan = _plan_from_cu_seqlens([int(v) for v in cu.cpu().tolist()], ctx.K)
            dqkv = _attention_backward(qkv2d, None, None, plan, ctx.K, ctx.H, ctx.scale, dout.to(qkv2d.dtype))
        return dqkv, None, None, None, None, None


def _arange(n, device):
    return torch.arange(n, dtype=torch.int64, device=device)


def flash_attn_varlen_qkvpacked_func(qkv, cu_seqlens, max_seqlen, dropout_p=0.0, softmax_scale=None, causal=False,
                                     window_size=(-1, -1), softcap=0.0, alibi_slopes=None, deterministic=False,
                                     return_attn_probs=False):
    """qkv [T, 3, H, d] (fp16 / bf16), cu_seqlens [B + 1] int32 -> out [T, H, d] in qkv's dtype.
    Every sequence is one attention patch (tcgen05 kernel: d in {16, 32, 48}, max_seqlen <= 1024; other shapes take
    the SIMT kernel).  fp16 inputs are computed in bf16 (the build's tensor-core type)."""
    if causal or alibi_slopes is not None or softcap != 0.0 or tuple(window_size) != (-1, -1) or return_attn_probs:
        raise NotImplementedError("flash_attn stand-in: only the plain varlen self-attention the reference calls")
    if dropout_p != 0.0:
        raise NotImplementedError("flash_attn stand-in: attention dropout is not supported (the lang configs use 0.0)")
    if qkv.dim() != 4 or qkv.shape[1] != 3:
        raise ValueError("qkv must be [total_tokens, 3, heads, head_dim]")
    T, _, H, d = qkv.shape
    scale = float(softmax_scale) if softmax_scale is not None else d ** -0.5
    x = qkv.reshape(T, 3 * H * d)
    x = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
    cu = cu_seqlens.to(torch.int32)
    table = torch.stack([cu[:-1], cu[1:], cu[:-1], cu[1:]], dim=1).contiguous()  # (q_begin, q_end, kv_begin, kv_end)
    K = int(max_seqlen)
    if torch.is_grad_enabled() and qkv.requires_grad:
        out = _VarlenAttnFn.apply(x.contiguous(), table, cu, K, H, scale)
    else:
        out = ops.patch_attention(x.contiguous(), _arange(T, x.device), table, K, H, scale)
    return out.reshape(T, H, d).to(qkv.dtype)


# ------------------------------------------------------------------------------------------- installation
def install(force: bool = False):
    """Register the stand-ins under the third-party module names.  Existing real packages are left alone unless
    `force` (then the reference runs on this package's kernels even where spconv / flash-attn are installed)."""
    def put(name, mod):
        if not force:
            try:
                __import__(name)          # a real installation wins
                return sys.modules[name]
            except Exception:
                pass
        sys.modules[name] = mod
        return mod

    ts = types.ModuleType("torch_scatter")
    ts.segment_csr = segment_csr
    fa = types.ModuleType("flash_attn")
    fa.flash_attn_varlen_qkvpacked_func = flash_attn_varlen_qkvpacked_func
    sp = types.ModuleType("spconv")
    spt = types.ModuleType("spconv.pytorch")
    spm = types.ModuleType("spconv.pytorch.modules")
    spt.SubMConv3d, spt.SparseConvTensor = spconv_compat.SubMConv3d, spconv_compat.SparseConvTensor
    spm.is_spconv_module = spconv_compat.is_spconv_module
    spt.modules, sp.pytorch = spm, spt
    for m in (ts, fa, sp, spt, spm):
        m.__scenesplat_b200__ = True
    out = {}
    for name, mod in (("torch_scatter", ts), ("flash_attn", fa), ("spconv", sp), ("spconv.pytorch", spt),
                      ("spconv.pytorch.modules", spm)):
        out[name] = put(name, mod)
    return out
