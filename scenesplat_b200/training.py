"""Training path of PT-v3m1 / LangPretrainer: forward AND backward.

The reference trains with torch autograd around three third-party operators (spconv SubMConv3d, flash-attn varlen,
torch_scatter segment_csr; pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py).  This module keeps
that structure: Linear / LayerNorm / BatchNorm(train) / GELU / DropPath / residuals are torch operators under autograd
(bf16 GEMMs through cuBLASLt, like the reference's AMP linears), and the three third-party operators are replaced by
`torch.autograd.Function`s over this package's kernels:

  SubMConvFn      forward: tcgen05 gather-GEMM + gather-sum (csrc/conv_gemm2.cu, conv_gemm.cu)
                  dgrad:   THE SAME kernels on the mirrored taps with transposed weights (the submanifold kernel map is
                           symmetric: nbr[t][p] = q  <=>  nbr[26-t][q] = p)
                  wgrad:   gathered-A x gathered-B tcgen05 GEMM per tap, split-K over pair chunks (csrc/conv_wgrad.cu)
  StemConvFn      forward: csrc/conv_simt.cu; wgrad per tap (the stem's input needs no gradient)
  PatchAttentionFn forward: tcgen05 patch attention (csrc/attention_tc.cu)
                  backward: recomputation with torch's fused SDPA on the gathered patches       [library, round 2:
                           tcgen05 backward kernel]
  SegmentMeanFn   forward: csrc/pool.cu segment_reduce; backward: gather / count

Indices (serialization, pooling levels, kernel maps, patch tables) carry no gradient and come from the same
`PointTransformerV3.prepare` as inference.  Gradients are checked against torch autograd through the CPU oracle
(tests/test_gpu_train.py).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .spconv_compat import kernel_map_for
from .structure import Dict, Point

BF16 = torch.bfloat16


# ------------------------------------------------------------------------------------------------ conv
def _pair_out(pairs):
    """Output voxel of every product row (inverse of ypos); padding rows point at voxel 0 and are never used.
    Built with one scatter (no nonzero / boolean indexing: those would drain the stream in the middle of the backward)."""
    if "pair_out" not in pairs:
        ypos = pairs["ypos"]                      # [k3, n] int32, -1 where the tap has no neighbour
        k3, n = ypos.shape
        p_pad = max(pairs["p_pad"], 1)
        po = torch.zeros(p_pad + 1, dtype=torch.long, device=ypos.device)   # slot p_pad swallows the missing taps
        idx = torch.where(ypos >= 0, ypos, torch.full_like(ypos, p_pad)).long().reshape(-1)
        po.scatter_(0, idx, torch.arange(n, device=ypos.device).repeat(k3))
        pairs["pair_out"] = po[:p_pad].contiguous()
    return pairs["pair_out"]


class SubMConvFn(torch.autograd.Function):
    """y[p] = b + sum_t W_t x[nbr[t][p]]  (3^3 submanifold conv, bf16 operands, fp32 accumulate)."""

    @staticmethod
    def forward(ctx, x, weight, bias, pairs, n):
        cout, cin = weight.shape[0], weight.shape[-1]
        k3 = weight.numel() // (cout * cin)
        w = weight.detach().reshape(cout, k3, cin).permute(1, 0, 2).contiguous().to(BF16)  # [k3, cout, cin]
        y = ops.subm_conv_gemm(x, pairs, w, bias.detach().float() if bias is not None else None, n, out_dtype=BF16)
        ctx.save_for_backward(x, weight)
        ctx.pairs, ctx.n, ctx.has_bias = pairs, n, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        pairs, n = ctx.pairs, ctx.n
        cout, cin = weight.shape[0], weight.shape[-1]
        k3 = weight.numel() // (cout * cin)
        dy = dy.contiguous().to(BF16)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # dx[q] = sum_t W_t^T dy[p : nbr[t][p] = q] = conv over the mirrored taps with W'_t = W_{k3-1-t}^T
            wt = weight.detach().reshape(cout, k3, cin).permute(1, 2, 0).flip(0).contiguous().to(BF16)  # [k3, cin, cout]
            dx = ops.subm_conv_gemm(dy, pairs, wt, None, n, out_dtype=BF16)
        if ctx.needs_input_grad[1]:
            # dW_t = dY[pair_out(t)]^T X[pair_in(t)]: gathered-operand GEMM on the tensor cores (csrc/conv_wgrad.cu)
            dw = ops.subm_conv_wgrad(x, dy, pairs, _pair_out(pairs), k3)
            dw = dw.permute(1, 0, 2).reshape(weight.shape).to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.float().sum(0)
        return dx, dw, db, None, None


class StemConvFn(torch.autograd.Function):
    """Stem conv (5^3, tiny Cin): SIMT forward, weight gradient per tap; the input carries no gradient."""

    @staticmethod
    def forward(ctx, x, weight, nbr):
        cout, cin = weight.shape[0], weight.shape[-1]
        k3 = weight.numel() // (cout * cin)
        wt = weight.detach().reshape(cout, k3, cin).permute(1, 2, 0).contiguous().float()  # [k3, cin, cout]
        y = ops.subm_conv_simt(x, nbr, wt, None, out_dtype=torch.float32)
        ctx.save_for_backward(x, weight, nbr)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, nbr = ctx.saved_tensors
        cout, cin = weight.shape[0], weight.shape[-1]
        k3 = weight.numel() // (cout * cin)
        dy = dy.float()
        n = x.shape[0]
        if cout == 32 and k3 * cin <= 1536 and x.dtype == torch.float32:  # csrc/conv_simt.cu: stem_wgrad kernels
            dw = ops.stem_conv_wgrad(x, dy, nbr, k3)                        # [k3, cin, cout]
            return None, dw.permute(2, 0, 1).reshape(weight.shape).to(weight.dtype), None
        xp = torch.cat([x.float(), x.new_zeros(1, cin, dtype=torch.float32)], 0)
        idx = torch.where(nbr >= 0, nbr, torch.full_like(nbr, n)).long()  # [k3, n]; missing neighbour -> the zero row
        dw = torch.matmul(xp[idx].transpose(1, 2), dy)                      # [k3, cin, n] @ [n, cout]
        return None, dw.permute(2, 0, 1).reshape(weight.shape).to(weight.dtype), None


# ------------------------------------------------------------------------------------------------ attention
def _patch_plan(point, K):
    """Patch layout of every batch item in SORTED positions (ref :114-170), from the offsets on the host:
    [(start, n_full_patches, tail_rows, end)], cached on the Point."""
    cache = point.setdefault("_patch_plan", {})
    if K not in cache:
        off = [0] + [int(v) for v in point.offset.cpu().tolist()]
        plan = []
        for s0, e0 in zip(off[:-1], off[1:]):
            nb = e0 - s0
            if nb <= 0:
                continue
            if nb <= K:
                plan.append((s0, 0, nb, e0))       # one short sequence: queries = keys = the whole item
            else:
                plan.append((s0, nb // K, nb - (nb // K) * K, e0))
        cache[K] = plan
    return cache[K]


def _heads(t, H):
    P, Lq, C = t.shape
    return t.reshape(P, Lq, H, C // H).transpose(1, 2)


def _attention_backward(qkv, order_row, inverse_row, plan, K, H, scale, dout):
    """d(qkv) of the patch attention, by recomputation with torch's fused attention (library; forward is the
    package's tcgen05 kernel).  Everything happens in SORTED space, where a patch is a contiguous row range: the
    rows are permuted in and out with the package's gather kernel (order / inverse are permutations, so both
    directions are gathers), full patches are one batched call on a view, the last patch of an item attends to
    the window of its last K rows and keeps its own tail rows."""
    C = qkv.shape[1] // 3
    if order_row is None:                        # rows already in sequence order (the flash-attn stand-in)
        xs, dos = qkv, dout.contiguous()
    else:
        xs = ops.gather_rows(qkv, order_row)    # [n, 3C] sorted
        dos = ops.gather_rows(dout.contiguous(), order_row)
    dxs = torch.zeros_like(xs)

    def run(q, k, v, do):
        q, k, v = (t.detach().requires_grad_(True) for t in (q, k, v))
        with torch.enable_grad():
            o = F.scaled_dot_product_attention(_heads(q, H), _heads(k, H), _heads(v, H), scale=scale)
        return torch.autograd.grad(o, (q, k, v), _heads(do, H))

    mains = [(s0, nf) for s0, nf, _, _ in plan if nf > 0]
    if mains:
        xm = [xs[s0:s0 + nf * K].view(nf, K, 3 * C) for s0, nf in mains]
        dm = [dos[s0:s0 + nf * K].view(nf, K, C) for s0, nf in mains]
        xm = torch.cat(xm, 0) if len(xm) > 1 else xm[0]
        dm = torch.cat(dm, 0) if len(dm) > 1 else dm[0]
        dq, dk, dv = run(xm[..., :C], xm[..., C:2 * C], xm[..., 2 * C:], dm)
        g = torch.cat([dq, dk, dv], -1)
        pos = 0
        for s0, nf in mains:
            dxs[s0:s0 + nf * K] = g[pos:pos + nf].reshape(nf * K, 3 * C)
            pos += nf
    for s0, nf, r, e0 in plan:
        if r > 0:
            w0 = max(e0 - K, s0)
            kv = xs[w0:e0].unsqueeze(0)
            q = xs[e0 - r:e0].unsqueeze(0)
            dq, dk, dv = run(q[..., :C], kv[..., C:2 * C], kv[..., 2 * C:], dos[e0 - r:e0].unsqueeze(0))
            dxs[e0 - r:e0, :C] += dq[0]
            dxs[w0:e0, C:2 * C] += dk[0]
            dxs[w0:e0, 2 * C:] += dv[0]
    return dxs if order_row is None else ops.gather_rows(dxs, inverse_row)


ATTN_BACKWARD = os.environ.get("SS_ATTN_BWD", "own")  # developer A/B switch: "sdpa" = library recomputation


def _own_attention_backward(head_dim, patch_size):
    return ATTN_BACKWARD == "own" and head_dim in (16, 32, 48) and patch_size <= 1024


class PatchAttentionFn(torch.autograd.Function):
    """Patch attention with the package's tcgen05 kernels in both directions (csrc/attention_tc.cu forward, which
    also emits the log-sum-exp of every query; csrc/attention_bwd.cu backward)."""

    @staticmethod
    def forward(ctx, qkv, order_row, inverse_row, table, K, H, scale, plan):
        own = plan is None
        if own:
            out, lse2 = ops.patch_attention_lse(qkv, order_row, table, K, H, scale)
            ctx.save_for_backward(qkv, order_row, inverse_row, table, out, lse2)
        else:
            out = ops.patch_attention(qkv, order_row, table, K, H, scale)
            ctx.save_for_backward(qkv, order_row, inverse_row)
        ctx.own, ctx.plan, ctx.K, ctx.H, ctx.scale = own, plan, K, H, scale
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.own:
            qkv, order_row, inverse_row, table, out, lse2 = ctx.saved_tensors
            dqkv = ops.patch_attention_backward(qkv, out, dout.to(BF16), lse2, order_row, table, ctx.K, ctx.H, ctx.scale)
        else:
            qkv, order_row, inverse_row = ctx.saved_tensors[:3]
            dqkv = _attention_backward(qkv, order_row, inverse_row, ctx.plan, ctx.K, ctx.H, ctx.scale, dout.to(qkv.dtype))
        return dqkv, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------ pooling
class SegmentMeanFn(torch.autograd.Function):
    """out[m] = mean_{j in [seg_start[m], seg_start[m+1])} src[order0[j]]   (replaces torch_scatter.segment_csr)."""

    @staticmethod
    def forward(ctx, src, order0, seg_start, cluster):
        out = ops.segment_reduce(src, order0, seg_start, "mean", out_dtype=torch.float32)
        ctx.save_for_backward(seg_start, cluster)
        ctx.dtype = src.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        seg_start, cluster = ctx.saved_tensors
        # dsrc[p] = dout[cluster[p]] / count[cluster[p]]: one gather kernel (csrc/pool_loss_bwd.cu)
        return ops.segment_mean_backward(dout, cluster, seg_start, "mean", out_dtype=ctx.dtype), None, None, None


class UnpoolGatherAddFn(torch.autograd.Function):
    """out = a + child[cluster]  (ref :478, `parent.feat + point.feat[inverse]`).  Backward: da = dout, dchild = the
    deterministic segment sum of dout over each cluster's members (torch's index backward is an atomic index_add)."""

    @staticmethod
    def forward(ctx, a, child, cluster, order0, seg_start):
        out, _ = ops.unpool_gather_add(a.contiguous(), child.contiguous().to(a.dtype), cluster, out_dtype=a.dtype)
        ctx.save_for_backward(order0, seg_start)
        ctx.child_dtype = child.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        order0, seg_start = ctx.saved_tensors
        return dout, ops.unpool_gather_add_backward(dout, order0, seg_start, out_dtype=ctx.child_dtype), None, None, None


class CosL2LossFn(torch.autograd.Function):
    """w_cos * mean_valid(1 - cos(pred, target)) + w_l2 * mean_valid ||pred - target||^2 (losses/misc.py:254-295) with the
    fused forward (ss_cos_l2_loss) and its one-pass adjoint (ss_cos_l2_loss_bwd); the valid count never leaves the device."""

    @staticmethod
    def forward(ctx, pred, target, mask, w_cos, w_l2):
        acc = ops.cos_l2_sums(pred, target, mask)
        ctx.save_for_backward(pred, target, mask, acc)
        ctx.w = (float(w_cos), float(w_l2))
        n_valid = acc[2].clamp(min=1.0)
        return ((w_cos * acc[0] + w_l2 * acc[1]) / n_valid).float()

    @staticmethod
    def backward(ctx, g):
        pred, target, mask, acc = ctx.saved_tensors
        return ops.cos_l2_backward(pred, target, mask, acc, g, *ctx.w).to(pred.dtype), None, None, None, None


class ClassHalfSumsFn(torch.autograd.Function):
    """Per (class, half) feature sums of AggregatedContrastiveLoss (misc.py:384-385) and their adjoint (a row gather)."""

    @staticmethod
    def forward(ctx, pred, valid, segment, half, n_classes):
        sums, counts = ops.class_half_sums(pred, valid, segment, half, n_classes)
        ctx.save_for_backward(valid, segment, half)
        ctx.n_classes, ctx.dtype = n_classes, pred.dtype
        ctx.mark_non_differentiable(counts)
        return sums, counts

    @staticmethod
    def backward(ctx, dsums, _dcounts):
        valid, segment, half = ctx.saved_tensors
        return ops.class_half_sums_backward(dsums, valid, segment, half, ctx.n_classes).to(ctx.dtype), None, None, None, None


# ------------------------------------------------------------------------------------------------ model walk
def _bf16_param(p):
    """bf16 copy of a parameter, made once per optimiser step (refreshed when the tensor's version counter or storage
    changes).  Kept on the parameter object itself, so it dies with the model (no process-global table)."""
    ent = p.__dict__.get("_ss_bf16")
    ver = (p._version, p.data_ptr())
    if ent is None or ent[0] != ver:
        ent = (ver, p.detach().to(BF16))
        p.__dict__["_ss_bf16"] = ent
    return ent[1]


class LinearFn(torch.autograd.Function):
    """nn.Linear on bf16 operands (the reference's AMP linears): cuBLASLt GEMMs as in the reference, but with the bf16
    parameter copies cached per step and the bias gradient from the package's column-sum kernel: 1 + 3 operator
    dispatches instead of torch's 4 + 9 (three casts, addmm, two mm, a strided reduction, three cast adjoints)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        w16 = _bf16_param(weight)
        y = F.linear(x, w16, _bf16_param(bias) if bias is not None else None)
        ctx.save_for_backward(x, w16)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w16 = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dy @ w16 if ctx.needs_input_grad[0] else None
        dw = dy.t() @ x if ctx.needs_input_grad[1] else None          # bf16 [Cout, Cin]; autograd casts to the parameter dtype
        db = ops.colsum(dy) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db


def _lin(m: nn.Linear, x):
    x = x.to(BF16)
    if x.is_cuda and x.dim() == 2 and m.out_features % 8 == 0:
        return LinearFn.apply(x.contiguous(), m.weight, m.bias)
    return F.linear(x, m.weight.to(BF16), m.bias.to(BF16) if m.bias is not None else None)


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over the channel dim: forward = the fused row kernel of the inference path (fp32 or bf16 in,
    fp32 or bf16 out, no separate cast passes), backward = csrc/backward.cu (one pass over x and dy)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype):
        w, b = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        x = x.contiguous()
        if x.dtype == torch.float32:
            _, y = ops.add_layernorm(x, None, None, (w, b), eps, want_res=False, norm_dtype=out_dtype)
        elif out_dtype == torch.float32:
            y, _ = ops.add_layernorm(None, x, (w, b), None, eps, want_res=True)
        else:
            _, y = ops.add_layernorm(None, x, (w, b), None, eps, want_res=False, norm_dtype=out_dtype)
        ctx.save_for_backward(x, w)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx, dg, db = ops.layernorm_backward(x, dy, w, ctx.eps)
        return dx, dg, db, None, None


class GeluFn(torch.autograd.Function):
    """nn.GELU() on the bf16 MLP hidden: the inference path's kernel forward, csrc/backward.cu backward."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.affine_act(x, act=1)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.gelu_backward(x, dy.to(BF16))


def _gelu(x):
    if x.dtype == BF16 and x.numel() % 8 == 0:
        return GeluFn.apply(x.contiguous())
    return F.gelu(x)


def _ln(m: nn.LayerNorm, x, out_dtype=torch.float32):
    c = x.shape[1]
    if c % 8 == 0 and c <= 1024 and x.dtype in (torch.float32, BF16):
        return LayerNormFn.apply(x, m.weight, m.bias, m.eps, out_dtype)
    return F.layer_norm(x.float(), m.normalized_shape, m.weight, m.bias, m.eps).to(out_dtype)


def _bn(m, x):
    """BatchNorm1d in the module's own mode (batch statistics + running-stat update when training)."""
    return m(x.float()) if m is not None else x


def _drop_path(seq, x):
    for m in seq.modules():
        if type(m).__name__ == "DropPath":
            return m(x)
    return x


def _pairs(point):
    ent = kernel_map_for(point, 3, want_pairs=True)
    return ent["pairs"]


def block_train(blk, point, x, conv_src=None):
    """Block.forward (ref :318-338) under autograd.  x: fp32 residual stream [N, C]."""
    conv, cpe_lin, cpe_ln = blk.cpe[0], blk.cpe[1], blk.cpe[2]
    src = x if conv_src is None else conv_src
    y = SubMConvFn.apply(src.to(BF16), conv.weight, conv.bias, _pairs(point), x.shape[0])
    x = x + _ln(cpe_ln, _lin(cpe_lin, y))
    att = blk.attn
    h = _ln(blk.norm1[0], x, BF16)
    qkv = _lin(att.qkv, h)
    table = att.patch_table(point)
    order_row = point.serialized_order[att.order_index].contiguous()
    inverse_row = point.serialized_inverse[att.order_index].contiguous()
    own = _own_attention_backward(qkv.shape[1] // (3 * att.num_heads), att.patch_size)
    plan = None if own else _patch_plan(point, att.patch_size)  # (the library path needs a host copy of the offsets)
    a = PatchAttentionFn.apply(qkv, order_row, inverse_row, table, att.patch_size, att.num_heads, att.scale, plan)
    x = x + _drop_path(blk.drop_path, _lin(att.proj, a))
    h = _ln(blk.norm2[0], x, BF16)
    mlp = blk.mlp[0]
    m = _lin(mlp.fc2, _gelu(_lin(mlp.fc1, h)))
    x = x + _drop_path(blk.drop_path, m)  # bf16 + fp32 -> fp32 inside the add (no separate cast pass)
    return x


def pooling_train(down, point, x):
    """SerializedPooling.forward (ref :371-444) under autograd; indices from the plan built by `prepare`."""
    plan = point["_pool_plan"]
    perm, ix = plan["perm"], plan["ix"]
    if down.reduce != "mean":
        raise NotImplementedError("training path: SerializedPooling(reduce='mean') only (the lang configs)")
    order0 = point.serialized_order[0].contiguous()
    feat = SegmentMeanFn.apply(_lin(down.proj, x), order0, ix["seg_start"], ix["cluster"])
    coord = ops.segment_reduce(point.coord.float(), order0, ix["seg_start"], "mean")
    names = point.serialized_order_names
    n_batch = point.offset.numel()
    offset = torch.searchsorted(ix["batch"], torch.arange(n_batch, device=feat.device), right=True)
    child = Point(Dict(feat=feat, coord=coord, grid_coord=ix["grid_coord"], serialized_code=ix["code"],
                       serialized_order=ix["order"], serialized_inverse=ix["inverse"],
                       serialized_depth=point.serialized_depth - plan["pooling_depth"], batch=ix["batch"], offset=offset,
                       serialized_order_names=tuple(names[i] for i in perm)))
    child["_kmap_cache"] = plan["kmap_cache"]
    if plan.get("child") is not None:
        child["_pool_plan"] = plan["child"]
    if getattr(down, "norm", None) is not None:
        feat = _bn(down.norm[0], feat)
    if getattr(down, "act", None) is not None:
        feat = F.gelu(feat)
    return child, feat, (ix["cluster"], order0, ix["seg_start"])


def unpool_train(up, x_child, pool_ix, x_parent):
    """SerializedUnpooling.forward (ref :471-482) under autograd -> (parent feat, skip branch alone).  pool_ix = the
    (cluster, order0, seg_start) of the pooling this level came from."""
    cluster, order0, seg_start = pool_ix
    lin_p, bn_p, act_p, _ = up._branch(up.proj)
    lin_s, bn_s, act_s, _ = up._branch(up.proj_skip)
    a = _bn(bn_p, _lin(lin_p, x_child).float())
    s = _bn(bn_s, _lin(lin_s, x_parent).float())
    if act_p is not None:
        a = F.gelu(a)
    if act_s is not None:
        s = F.gelu(s)
    return UnpoolGatherAddFn.apply(s, a, cluster, order0, seg_start), s


def forward_train(model, data_dict):
    """PointTransformerV3.forward (ref :699-714) under autograd -> fp32 features [N, C_dec0]."""
    point = model.prepare(data_dict)
    stem = model.embedding.stem._modules
    ent = kernel_map_for(point, stem["conv"].kernel_size, want_pairs=False)
    x = StemConvFn.apply(point.feat.float().contiguous(), stem["conv"].weight, ent["nbr"])
    if "norm" in stem:
        x = _bn(stem["norm"], x)
    if "act" in stem:
        x = F.gelu(x)
    skips = []
    for stage in model.enc.children():
        down = getattr(stage, "down", None)
        if down is not None:
            skips.append((point, x))
            point, x, cluster = pooling_train(down, point, x)
            skips[-1] = skips[-1] + (cluster,)
        for name, blk in stage.named_children():
            if name != "down":
                x = block_train(blk, point, x)
    if model.cls_mode:
        return x
    for stage in model.dec.children():
        parent, x_parent, cluster = skips.pop()
        x, skip_only = unpool_train(stage.up, x, cluster, x_parent)
        point = parent
        first = True
        for name, blk in stage.named_children():
            if name != "up":
                # ref quirk: after unpooling, sparse_conv_feat still holds the skip projection alone (see ptv3.py)
                x = block_train(blk, point, x, conv_src=skip_only if first else None)
                first = False
    return x


# ------------------------------------------------------------------------------------------------ losses (autograd)
def _fused_loss_ok(pred, target):
    return pred.is_cuda and pred.dim() == 2 and pred.shape[1] % 8 == 0 and pred.dtype in (torch.float32, BF16) and \
        target.dtype in (torch.float32, BF16, torch.float16)


def cosine_loss(pred, target, mask, loss_weight=1.0):
    """losses/misc.py:254-270: mean over the valid rows of 1 - cos: fused forward + adjoint kernels (no boolean
    indexing, no host decision; no valid row -> 0).  Shapes the kernels do not take use the torch formulation below."""
    if _fused_loss_ok(pred, target):
        return CosL2LossFn.apply(pred, target, mask.reshape(-1), float(loss_weight), 0.0)
    m = mask.reshape(-1).bool()
    cos = F.cosine_similarity(pred.float(), target.float(), dim=1, eps=1e-8)
    n_valid = m.sum().clamp(min=1).to(cos.dtype)
    return loss_weight * (torch.where(m, 1.0 - cos, torch.zeros_like(cos)).sum() / n_valid)


def l2_loss(pred, target, mask, loss_weight=1.0):
    """losses/misc.py:280-295: mean over the valid rows of the squared distance (same kernels as cosine_loss)."""
    if _fused_loss_ok(pred, target):
        return CosL2LossFn.apply(pred, target, mask.reshape(-1), 0.0, float(loss_weight))
    m = mask.reshape(-1).bool()
    d2 = ((pred.float() - target.float()) ** 2).sum(1)
    n_valid = m.sum().clamp(min=1).to(d2.dtype)
    return loss_weight * (torch.where(m, d2, torch.zeros_like(d2)).sum() / n_valid)


def contrastive_from_sums(sums, counts, n_classes, temperature, reduction="mean"):
    """losses/misc.py:376-412 on the per-(class, half) sums, over a FIXED number of class slots: classes that do not
    qualify (fewer than 100 points or an empty half) are masked out of the logits and of the mean instead of being
    removed with nonzero(), so the loss needs no host decision (no qualifying class -> 0)."""
    c2 = counts.view(n_classes, 2)
    use = (c2.sum(1) >= 100) & (c2.min(1).values > 0)
    s = sums.view(n_classes, 2, -1).float()
    a = F.normalize(s[:, 0], p=2, dim=1)
    b = F.normalize(s[:, 1], p=2, dim=1)
    logits = (a @ b.t()) / temperature
    neg = torch.full_like(logits, -1e30)        # finite: an all-masked row still has a defined softmax gradient
    n_used = use.sum().clamp(min=1).to(logits.dtype)
    diag = logits.diagonal()

    def ce(lg):
        lse = torch.logsumexp(torch.where(use[None, :], lg, neg), dim=1)
        return torch.where(use, lse - diag, torch.zeros_like(diag)).sum() / n_used

    loss = (ce(logits) + ce(logits.t())) / 2.0
    if reduction == "sum":
        loss = loss * use.sum().to(logits.dtype)
    return loss


def class_half_sums(pred, valid, segment, half, n_classes):
    """Per (class, half) feature sums under autograd: the fused kernel + its gather adjoint; torch index_add for shapes
    the kernel does not take."""
    if pred.is_cuda and pred.shape[1] % 8 == 0 and pred.dtype in (torch.float32, BF16):
        sums, counts = ClassHalfSumsFn.apply(pred, valid, segment, half, n_classes)
        return sums, counts.long()
    key = segment.long().clamp(min=0) * 2 + half.long()
    w = valid.to(pred.dtype)[:, None]
    sums = pred.new_zeros(2 * n_classes, pred.shape[1]).index_add(0, key, pred * w)
    counts = torch.zeros(2 * n_classes, dtype=torch.long, device=pred.device).index_add(0, key, valid.long())
    return sums, counts
