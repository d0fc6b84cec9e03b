"""Point Transformer V3 (mode 1) on the sm_100a kernels -- drop-in for
pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py: same class names, constructor
signatures, sub-module names and `state_dict` keys/shapes (393 entries for the lang config), same
`Point` dict keys in and out, same CPU-RNG consumption (`torch.randperm`) so seeded runs shuffle the
serialization orders identically.

What changes underneath (eval / inference path):
  * Block (ref :318-338): xCPE conv = kernel map from the sorted keys + tcgen05 gather-GEMM; the three
    residual adds and four LayerNorms collapse into three fused row kernels; attention reads q/k/v
    through the serialized order and writes back through it (no [order] / [inverse] gather passes).
  * SerializedPooling (ref :371-444): cluster ids / counts / pooled code, order, inverse of all rows come
    from run-length scans along the parent's orders (no unique, no sort), segment mean + BN + GELU fused.
  * SerializedUnpooling (ref :471-482): gather-add with both BN+GELU branches fused.
  * MLP (ref :225-248): fc1 + bias + exact GELU is ONE tcgen05 kernel on CTA pairs (csrc/gemm2cta.cu); the xCPE Linear
    is folded into the conv's per-tap weights.  The remaining dense Linear layers are library GEMMs (cuBLASLt
    through torch), in bf16 with fp32 accumulation.
Training (`model.train()` under autograd) takes scenesplat_b200/training.py.
"""
from __future__ import annotations

import math
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import ops
from . import spconv_compat as spconv
from .modules import PointModule, PointSequential, _apply
from .registry import MODELS
from .structure import Dict, Point, offset2bincount

BF16 = torch.bfloat16


class DropPath(nn.Module):
    """timm.layers.DropPath (per-row stochastic depth); identity in eval."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


def _no_training(mod):
    if torch.is_grad_enabled() and mod.training:
        raise NotImplementedError(
            f"scenesplat_b200: {type(mod).__name__}.forward is the inference path (fused, no autograd graph).  "
            "Training runs through PointTransformerV3.forward / LangPretrainer.forward, which route to "
            "scenesplat_b200/training.py; to call a sub-module on its own use .eval() and torch.no_grad()")


class _Cache:
    """bf16 / folded copies of parameters, refreshed when a parameter is modified in place or rebound (load_state_dict,
    `.to()`).  The copies live ON the owning module (a plain attribute, not a buffer: never in `state_dict`), so they
    are freed with the model and can never be served to another one."""

    ATTR = "_ss_prepared"

    def get(self, owner, key, params, fn):
        store = owner.__dict__.get(self.ATTR)
        if store is None:
            store = {}
            object.__setattr__(owner, self.ATTR, store)
        ver = tuple((p._version, p.data_ptr()) for p in params)
        ent = store.get(key)
        if ent is None or ent[0] != ver:
            with torch.no_grad():
                ent = (ver, fn())
            store[key] = ent
        return ent[1]


_cache = _Cache()


def linear_params(lin: nn.Linear):
    """(bf16 weight [cout, cin], fp32 bias or None) of a Linear, cached on the module."""
    return _cache.get(lin, "lin", [lin.weight] + ([lin.bias] if lin.bias is not None else []),
                      lambda: (lin.weight.detach().to(BF16).contiguous(),
                               lin.bias.detach().float().clone() if lin.bias is not None else None))


def linear_bn_params(lin: nn.Linear, bn):
    """Linear followed by an eval-mode BatchNorm1d as ONE Linear: (bf16 weight * scale[:, None], fp32 bias * scale + shift),
    composed in fp32 and cached on the Linear (both modules' parameters and running statistics are the cache key)."""
    if bn is None:
        return linear_params(lin)

    def fn():
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
        shift = bn.bias.detach().float() - bn.running_mean.float() * scale
        w = (lin.weight.detach().float() * scale[:, None]).to(BF16).contiguous()
        b = shift if lin.bias is None else lin.bias.detach().float() * scale + shift
        return w, b.contiguous()
    deps = [lin.weight] + ([lin.bias] if lin.bias is not None else []) + [bn.weight, bn.bias, bn.running_mean, bn.running_var]
    return _cache.get(lin, "lin_bn", deps, fn)


def linear_bn_act_bf16(lin: nn.Linear, bn, x, act=0):
    """act(BatchNorm1d_eval(Linear(x))) in the GEMM's epilogue (BN folded into the weights, exact GELU when act = 1)."""
    w, b = linear_bn_params(lin, bn)
    if x.dtype != BF16:
        x = x.to(BF16)
    if x.is_cuda and ops.linear_ok(lin.in_features, lin.out_features):
        return ops.linear_act(x, w, b, act=act)
    y = F.linear(x, w, b.to(BF16) if b is not None else None)
    return ops.affine_act(y, act=1) if act else y


def linear_bf16(lin: nn.Linear, x, act=0):
    """nn.Linear (+ exact GELU when act = 1) on bf16 operands with fp32 accumulation: the package's tcgen05 CTA-pair GEMM
    (csrc/gemm2cta.cu) with bias / activation in the epilogue.  Channel counts outside its tiling (K not a multiple of
    16, output columns not a multiple of 32: none in the PTv3 configs) go to the library GEMM."""
    w, b = linear_params(lin)
    if x.dtype != BF16:
        x = x.to(BF16)
    if x.is_cuda and ops.linear_ok(lin.in_features, lin.out_features):
        return ops.linear_act(x, w, b, act=act)
    y = F.linear(x, w, b.to(BF16) if b is not None else None)
    return ops.affine_act(y, act=1) if act else y


def cpe_folded(conv, lin: nn.Linear):
    """xCPE conv followed by its Linear (ref :277-287) as ONE conv: Linear(b_c + sum_t W_t x_t) =
    (W_l b_c + b_l) + sum_t (W_l W_t) x_t.  The composed per-tap weights [k^3, C_out, C_in] are formed once in fp32
    (eval: the parameters are constants) and rounded to bf16; the C x C GEMM over all N voxels and the bf16 round trip
    of the conv output disappear from every Block."""
    def fn():
        cout, cin = conv.out_channels, conv.in_channels
        k3 = conv.kernel_size ** 3
        wc = conv.weight.detach().float().reshape(cout, k3, cin).permute(1, 0, 2)       # [k3, cout, cin]
        wl = lin.weight.detach().float()                                                # [C, cout]
        w = torch.matmul(wl.unsqueeze(0), wc).to(BF16).contiguous()                    # [k3, C, cin]
        b = lin.bias.detach().float() if lin.bias is not None else torch.zeros(lin.out_features, device=wl.device)
        if conv.bias is not None:
            b = b + wl @ conv.bias.detach().float()
        return w, b.contiguous()
    deps = [conv.weight, lin.weight] + [t for t in (conv.bias, lin.bias) if t is not None]
    return _cache.get(lin, "cpe", deps, fn)


def ln_params(ln: nn.LayerNorm):
    return _cache.get(ln, "ln", [ln.weight, ln.bias],
                      lambda: (ln.weight.detach().float().clone(), ln.bias.detach().float().clone()))  # own, aligned storage


def bn_fold(bn: nn.BatchNorm1d):
    """eval-mode BatchNorm1d -> per-channel (scale, shift)."""
    def fn():
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
        shift = bn.bias.detach().float() - bn.running_mean.float() * scale
        return scale.contiguous(), shift.contiguous()
    return _cache.get(bn, "bn", [bn.weight, bn.bias, bn.running_mean, bn.running_var], fn)


def _single(seq, cls):
    """The one `cls` module inside a PointSequential (or None)."""
    if seq is None:
        return None
    mods = [m for m in seq._modules.values() if isinstance(m, cls)]
    return mods[0] if len(mods) == 1 else None


def _bf16_of(point, feat):
    """bf16 copy of `feat`, reusing the shadow the previous fused kernel already wrote."""
    sh = point.get("_bf16_shadow")
    if sh is not None and sh[0] is feat:
        return sh[1]
    return feat if feat.dtype == BF16 else feat.to(BF16)


class RPE(torch.nn.Module):
    def __init__(self, patch_size, num_heads):
        super().__init__()
        self.patch_size, self.num_heads = patch_size, num_heads
        self.pos_bnd = int((4 * patch_size) ** (1 / 3) * 2)
        self.rpe_num = 2 * self.pos_bnd + 1
        self.rpe_table = torch.nn.Parameter(torch.zeros(3 * self.rpe_num, num_heads))
        torch.nn.init.trunc_normal_(self.rpe_table, std=0.02)

    def forward(self, coord):
        raise NotImplementedError("relative position encoding is not on the flash path the lang configs use")


class SerializedAttention(PointModule):
    """ref :51-222"""

    def __init__(self, channels, num_heads, patch_size, qkv_bias=True, qk_scale=None, attn_drop=0.0, proj_drop=0.0,
                 order_index=0, enable_rpe=False, enable_flash=True, upcast_attention=True, upcast_softmax=True):
        super().__init__()
        assert channels % num_heads == 0
        self.channels, self.num_heads = channels, num_heads
        self.scale = qk_scale or (channels // num_heads) ** -0.5
        self.order_index = order_index
        self.upcast_attention, self.upcast_softmax = upcast_attention, upcast_softmax
        self.enable_rpe, self.enable_flash = enable_rpe, enable_flash
        if enable_flash:
            assert enable_rpe is False, "Set enable_rpe to False when enable Flash Attention"
            assert upcast_attention is False, "Set upcast_attention to False when enable Flash Attention"
            assert upcast_softmax is False, "Set upcast_softmax to False when enable Flash Attention"
            self.patch_size = patch_size
            self.attn_drop = attn_drop
        else:
            self.patch_size_max = patch_size
            self.patch_size = 0
            self.attn_drop = torch.nn.Dropout(attn_drop)
        self.qkv = torch.nn.Linear(channels, channels * 3, bias=qkv_bias)
        self.proj = torch.nn.Linear(channels, channels)
        self.proj_drop = torch.nn.Dropout(proj_drop)
        self.softmax = torch.nn.Softmax(dim=-1)
        self.rpe = RPE(patch_size, num_heads) if self.enable_rpe else None

    def patch_table(self, point):
        """Device-side patch table (q/kv ranges in sorted positions); replaces get_padding_and_inverse (:114-170)."""
        if not self.enable_flash:  # ref :173-176
            self.patch_size = min(int(offset2bincount(point.offset).min()), self.patch_size_max)
        tabs = point.setdefault("_patch_tables", {})
        K = self.patch_size
        if K not in tabs:
            tabs[K] = ops.patch_table(point.offset, K, point.feat.shape[0])
        return tabs[K]

    @torch.no_grad()
    def get_padding_and_inverse(self, point):
        """Materialise the reference's `pad` / `unpad` / `cu_seqlens_key` (only if a caller wants them)."""
        if not {"pad", "unpad", "cu_seqlens_key"}.issubset(point.keys()):
            t = self.patch_table(point)
            qb, qe, kb, ke = (t[:, i].long() for i in range(4))
            live = qe > qb
            qb, qe, kb, ke = qb[live], qe[live], kb[live], ke[live]
            lens = ke - kb
            cu = F.pad(torch.cumsum(lens, 0), (1, 0))
            pos = torch.arange(int(cu[-1]), device=t.device)
            seq = torch.searchsorted(cu[1:], pos, right=True)
            local = pos - cu[:-1][seq]
            own = (qe - qb)[seq]
            pad = torch.where(local < own, qb[seq] + local, kb[seq] + (local - own))
            unpad = torch.empty(point.feat.shape[0], dtype=torch.long, device=t.device)
            is_own = local < own
            unpad[pad[is_own]] = pos[is_own]
            point["pad"], point["unpad"], point["cu_seqlens_key"] = pad, unpad, cu.int()
        return point["pad"], point["unpad"], point["cu_seqlens_key"]

    def core(self, point, qkv):
        table = self.patch_table(point)
        return ops.patch_attention(qkv, point.serialized_order[self.order_index].contiguous(), table, self.patch_size,
                                   self.num_heads, self.scale)

    def forward(self, point):
        _no_training(self)
        if self.enable_rpe:
            raise NotImplementedError("enable_rpe")
        qkv = linear_bf16(self.qkv, point.feat)
        feat = self.core(point, qkv)
        point.feat = linear_bf16(self.proj, feat)
        return point


class MLP(nn.Module):
    """ref :225-248"""

    def __init__(self, in_channels, hidden_channels=None, out_channels=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_channels = out_channels or in_channels
        hidden_channels = hidden_channels or in_channels
        self.fc1 = nn.Linear(in_channels, hidden_channels)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_channels, out_channels)
        self.drop = nn.Dropout(drop)

    def _exact_gelu(self):
        return isinstance(self.act, nn.GELU) and self.act.approximate == "none"

    def forward(self, x):
        # fc1 + bias + GELU is ONE kernel (the N x 4C hidden is written once), fc2 + bias another
        if self._exact_gelu():
            h = linear_bf16(self.fc1, x, act=1)
        else:
            h = self.act(linear_bf16(self.fc1, x))
        return linear_bf16(self.fc2, h)

    def forward_residual(self, x, res):
        """res + MLP(x) with the residual add in fc2's epilogue (eval, DropPath = identity): -> (fp32 residual stream,
        updated in place, and its bf16 copy), or None when the shapes are outside the fused kernel."""
        if not (self._exact_gelu() and x.is_cuda and ops.linear_ok(self.fc2.in_features, self.fc2.out_features)):
            return None
        h = linear_bf16(self.fc1, x, act=1)
        w, b = linear_params(self.fc2)
        return ops.linear_residual(h, w, b, res)


class Block(PointModule):
    """ref :251-338"""

    def __init__(self, channels, num_heads, patch_size=48, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, attn_drop=0.0,
                 proj_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, act_layer=nn.GELU, pre_norm=True, order_index=0,
                 cpe_indice_key=None, enable_rpe=False, enable_flash=True, upcast_attention=True, upcast_softmax=True):
        super().__init__()
        self.channels, self.pre_norm = channels, pre_norm
        self.cpe = PointSequential(
            spconv.SubMConv3d(channels, channels, kernel_size=3, bias=True, indice_key=cpe_indice_key),
            nn.Linear(channels, channels),
            norm_layer(channels),
        )
        self.norm1 = PointSequential(norm_layer(channels))
        self.attn = SerializedAttention(channels=channels, patch_size=patch_size, num_heads=num_heads, qkv_bias=qkv_bias,
                                        qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=proj_drop,
                                        order_index=order_index, enable_rpe=enable_rpe, enable_flash=enable_flash,
                                        upcast_attention=upcast_attention, upcast_softmax=upcast_softmax)
        self.norm2 = PointSequential(norm_layer(channels))
        self.mlp = PointSequential(MLP(in_channels=channels, hidden_channels=int(channels * mlp_ratio),
                                       out_channels=channels, act_layer=act_layer, drop=proj_drop))
        self.drop_path = PointSequential(DropPath(drop_path) if drop_path > 0.0 else nn.Identity())

    def forward(self, point: Point):
        _no_training(self)
        if not (self.pre_norm and all(isinstance(m, nn.LayerNorm) for m in (self.cpe[2], self.norm1[0], self.norm2[0]))):
            raise NotImplementedError("only pre-norm Blocks with LayerNorm are built (the lang configs); PDNorm is off")
        conv, cpe_lin, cpe_ln = self.cpe[0], self.cpe[1], self.cpe[2]
        x = point.feat
        if x.dtype != torch.float32:
            x = x.float()
        # xCPE: the conv reads sparse_conv_feat.features (modules.py:68-72), which after an unpooling is the
        # skip projection only (reference quirk, see oracle/ptv3.py: unpooling_forward)
        src = point.sparse_conv_feat.features
        if conv.tensor_core_ok() and cpe_lin.out_features % 32 == 0:
            w_fold, b_fold = cpe_folded(conv, cpe_lin)  # conv and Linear composed into one tensor-core conv
            ent = spconv.kernel_map_for(point, conv.kernel_size, want_pairs=True)
            if self.channels >= 128 and cpe_ln.eps == self.norm1[0].eps:
                # gather-sum + LN(cpe) + residual add + LN(norm1) in one pass: the conv output stays in registers
                x, h = ops.subm_conv_gemm_add_ln(_bf16_of(point, src), ent["pairs"], w_fold, b_fold, x.contiguous(),
                                                 ln_params(cpe_ln), ln_params(self.norm1[0]), cpe_ln.eps, inplace=False)
                z = None
            else:
                z = ops.subm_conv_gemm(_bf16_of(point, src), ent["pairs"], w_fold, b_fold, x.shape[0], out_dtype=BF16)
        else:
            y = conv.conv_point(point, _bf16_of(point, src))
            z = linear_bf16(cpe_lin, y)
        if z is not None:
            x, h = ops.add_layernorm(x, z, ln_params(cpe_ln), ln_params(self.norm1[0]), cpe_ln.eps, norm_dtype=BF16)
        qkv = linear_bf16(self.attn.qkv, h)
        a = self.attn.core(point, qkv)
        p = linear_bf16(self.attn.proj, a)
        x, h = ops.add_layernorm(x, p, None, ln_params(self.norm2[0]), self.norm2[0].eps, norm_dtype=BF16, inplace=True)
        # (MLP.forward_residual -- fc2 with the residual add in its epilogue -- is built and tested but measured SLOWER
        # than GEMM + this add pass at every Block shape: 1.43 vs 1.37 ms at dec0, profiles/r2_gemm.md)
        m = self.mlp[0](h)
        if point.pop("_emit_l2_normalized", False) and x.shape[1] % 8 == 0 and x.shape[1] <= 1024:
            # last Block of the backbone under LangPretrainer (eval): residual add + F.normalize(p=2, dim=1) in one pass;
            # neither the un-normalised sum nor its bf16 copy is written (models/default.py:98)
            x = ops.add_l2_normalize(x, m, eps=1e-12)
            point.feat = x
            point["_l2_normalized"] = True
            point.sparse_conv_feat = point.sparse_conv_feat.replace_feature(x)
            return point
        x, xb = ops.add_layernorm(x, m, None, None, norm_dtype=BF16, inplace=True)
        point.feat = x
        point["_bf16_shadow"] = (x, xb)
        point.sparse_conv_feat = point.sparse_conv_feat.replace_feature(x)
        return point


class SerializedPooling(PointModule):
    """ref :341-444"""

    def __init__(self, in_channels, out_channels, stride=2, norm_layer=None, act_layer=None, reduce="mean",
                 shuffle_orders=True, traceable=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        assert stride == 2 ** (math.ceil(stride) - 1).bit_length()
        self.stride = stride
        assert reduce in ["sum", "mean", "min", "max"]
        self.reduce, self.shuffle_orders, self.traceable = reduce, shuffle_orders, traceable
        self.proj = nn.Linear(in_channels, out_channels)
        if norm_layer is not None:
            self.norm = PointSequential(norm_layer(out_channels))
        if act_layer is not None:
            self.act = PointSequential(act_layer())

    def pooling_depth_for(self, serialized_depth):
        pooling_depth = (math.ceil(self.stride) - 1).bit_length()
        return 0 if pooling_depth > serialized_depth else pooling_depth

    def pool_indices(self, point, pooling_depth):
        """ref :384-412: cluster ids, pooled codes / orders / inverses / grid_coord / batch (one host sync: the
        number of coarse voxels) and the order shuffle (one `torch.randperm` from the CPU generator)."""
        k = point.serialized_code.shape[0]
        perm = torch.randperm(k).tolist() if self.shuffle_orders else list(range(k))
        ix = ops.pool_index(point.serialized_code, point.serialized_order, point.grid_coord, point.batch, pooling_depth,
                            perm)
        return perm, ix

    def forward(self, point: Point):
        _no_training(self)
        pooling_depth = self.pooling_depth_for(point.serialized_depth)
        assert {"serialized_code", "serialized_order", "serialized_inverse", "serialized_depth"}.issubset(point.keys()), \
            "Run point.serialization() point cloud before SerializedPooling"
        plan = point.get("_pool_plan", None)  # index hierarchy built ahead of the feature path (PointTransformerV3.forward)
        if plan is not None and plan["pooling_depth"] == pooling_depth:
            perm, ix = plan["perm"], plan["ix"]
        else:
            plan = None
            perm, ix = self.pool_indices(point, pooling_depth)
        order0 = point.serialized_order[0].contiguous()
        bn = _single(self.norm, nn.BatchNorm1d) if getattr(self, "norm", None) is not None else None
        act = getattr(self, "act", None)
        gelu = act is not None and isinstance(act[0], nn.GELU) and act[0].approximate == "none"
        fuse = (bn is not None or getattr(self, "norm", None) is None) and (gelu or act is None)
        proj = linear_bf16(self.proj, _bf16_of(point, point.feat))
        scale, shift = bn_fold(bn) if (fuse and bn is not None) else (None, None)
        feat, coord = ops.pool_reduce(proj, point.coord.float(), order0, ix["seg_start"], self.reduce, scale, shift,
                                      1 if (fuse and gelu) else 0, out_dtype=torch.float32)
        n_batch = point.offset.numel()
        offset = torch.searchsorted(ix["batch"], torch.arange(n_batch, device=feat.device), right=True)
        names = point.serialized_order_names
        point_dict = Dict(
            feat=feat, coord=coord, grid_coord=ix["grid_coord"], serialized_code=ix["code"], serialized_order=ix["order"],
            serialized_inverse=ix["inverse"], serialized_depth=point.serialized_depth - pooling_depth, batch=ix["batch"],
            offset=offset, serialized_order_names=tuple(names[i] for i in perm),
        )
        if "condition" in point.keys():
            point_dict["condition"] = point.condition
        if "context" in point.keys():
            point_dict["context"] = point.context
        if self.traceable:
            point_dict["pooling_inverse"] = ix["cluster"]
            point_dict["pooling_parent"] = point
        point = Point(point_dict)
        if plan is not None:
            point["_kmap_cache"] = plan["kmap_cache"]
            if plan.get("child") is not None:
                point["_pool_plan"] = plan["child"]
        if not fuse:
            if getattr(self, "norm", None) is not None:
                point = self.norm(point)
            if act is not None:
                point = self.act(point)
        point.sparsify()
        return point


class SerializedUnpooling(PointModule):
    """ref :447-482"""

    def __init__(self, in_channels, skip_channels, out_channels, norm_layer=None, act_layer=None, traceable=False):
        super().__init__()
        self.proj = PointSequential(nn.Linear(in_channels, out_channels))
        self.proj_skip = PointSequential(nn.Linear(skip_channels, out_channels))
        if norm_layer is not None:
            self.proj.add(norm_layer(out_channels))
            self.proj_skip.add(norm_layer(out_channels))
        if act_layer is not None:
            self.proj.add(act_layer())
            self.proj_skip.add(act_layer())
        self.traceable = traceable

    @staticmethod
    def _branch(seq):
        mods = list(seq._modules.values())
        lin = mods[0]
        bn = next((m for m in mods[1:] if isinstance(m, nn.BatchNorm1d)), None)
        act = next((m for m in mods[1:] if isinstance(m, nn.GELU)), None)
        plain = len(mods) == 1 + (bn is not None) + (act is not None) and (act is None or act.approximate == "none")
        return lin, bn, act, plain

    def forward(self, point):
        _no_training(self)
        assert "pooling_parent" in point.keys()
        assert "pooling_inverse" in point.keys()
        parent = point.pop("pooling_parent")
        inverse = point.pop("pooling_inverse")
        lin_p, bn_p, act_p, ok_p = self._branch(self.proj)
        lin_s, bn_s, act_s, ok_s = self._branch(self.proj_skip)
        if not (ok_p and ok_s and (act_p is None) == (act_s is None)):
            raise NotImplementedError("SerializedUnpooling: only Linear[+BatchNorm1d][+GELU] branches are built")
        # Both branches are Linear -> BatchNorm1d(eval) -> GELU (ref :447-470): BN is folded into the Linear and the GELU runs
        # in the GEMM's epilogue on the fp32 accumulators, so the coarse branch is activated once per COARSE row (not once
        # per fine voxel that gathers it) and the pass below is a plain gather-add.  (ncu put the previous form, which
        # evaluated both activations inside the gather pass, at 52 % XU / 69 % issue: bound by the two GELUs per element.)
        act = 1 if act_s is not None else 0
        a = linear_bn_act_bf16(lin_p, bn_p, _bf16_of(point, point.feat), act)
        # `skip` is only ever read as the bf16 operand of the next block's xCPE conv: the GEMM's bf16 output is it
        skip = linear_bn_act_bf16(lin_s, bn_s, _bf16_of(parent, parent.feat), act)
        out, _ = ops.unpool_gather_add(skip, a, inverse, None, None, None, None, 0, out_dtype=torch.float32, want_a=False)
        parent.feat = out
        # ref :478 rebinds parent.feat only: sparse_conv_feat keeps the skip projection (oracle/ptv3.py)
        parent.sparse_conv_feat = parent.sparse_conv_feat.replace_feature(skip)
        parent.pop("_bf16_shadow", None)
        if self.traceable:
            parent["unpooling_parent"] = point
        return parent


class Embedding(PointModule):
    """ref :485-515"""

    def __init__(self, in_channels, embed_channels, norm_layer=None, act_layer=None):
        super().__init__()
        self.in_channels, self.embed_channels = in_channels, embed_channels
        self.stem = PointSequential(conv=spconv.SubMConv3d(in_channels, embed_channels, kernel_size=5, padding=1,
                                                           bias=False, indice_key="stem"))
        if norm_layer is not None:
            self.stem.add(norm_layer(embed_channels), name="norm")
        if act_layer is not None:
            self.stem.add(act_layer(), name="act")

    def forward(self, point: Point):
        _no_training(self)
        mods = self.stem._modules
        bn = mods.get("norm")
        act = mods.get("act")
        if (bn is None or isinstance(bn, nn.BatchNorm1d)) and (act is None or (isinstance(act, nn.GELU)
                                                                                 and act.approximate == "none")):
            scale, shift = bn_fold(bn) if bn is not None else (None, None)
            src = point.sparse_conv_feat.features
            feat = mods["conv"].conv_point(point, src.float() if src.dtype not in (torch.float32, BF16) else src, scale,
                                           shift, 1 if act is not None else 0, out_dtype=torch.float32)
            point.feat = feat
            point.sparse_conv_feat = point.sparse_conv_feat.replace_feature(feat)
            return point
        return self.stem(point)


@MODELS.register_module("PT-v3m1")
class PointTransformerV3(PointModule):
    """ref :518-714 (registered as "PT-v3m1")."""

    def __init__(self, in_channels=6, order=("z", "z-trans"), stride=(2, 2, 2, 2), enc_depths=(2, 2, 2, 6, 2),
                 enc_channels=(32, 64, 128, 256, 512), enc_num_head=(2, 4, 8, 16, 32),
                 enc_patch_size=(48, 48, 48, 48, 48), dec_depths=(2, 2, 2, 2), dec_channels=(64, 64, 128, 256),
                 dec_num_head=(4, 4, 8, 16), dec_patch_size=(48, 48, 48, 48), mlp_ratio=4, qkv_bias=True, qk_scale=None,
                 attn_drop=0.0, proj_drop=0.0, drop_path=0.3, pre_norm=True, shuffle_orders=True, enable_rpe=False,
                 enable_flash=True, upcast_attention=False, upcast_softmax=False, cls_mode=False, pdnorm_bn=False,
                 pdnorm_ln=False, pdnorm_decouple=True, pdnorm_adaptive=False, pdnorm_affine=True,
                 pdnorm_conditions=("ScanNet", "S3DIS", "Structured3D")):
        super().__init__()
        self.num_stages = len(enc_depths)
        self.order = [order] if isinstance(order, str) else order
        self.cls_mode = cls_mode
        self.shuffle_orders = shuffle_orders
        assert self.num_stages == len(stride) + 1
        assert self.num_stages == len(enc_depths)
        assert self.num_stages == len(enc_channels)
        assert self.num_stages == len(enc_num_head)
        assert self.num_stages == len(enc_patch_size)
        assert self.cls_mode or self.num_stages == len(dec_depths) + 1
        assert self.cls_mode or self.num_stages == len(dec_channels) + 1
        assert self.cls_mode or self.num_stages == len(dec_num_head) + 1
        assert self.cls_mode or self.num_stages == len(dec_patch_size) + 1
        if pdnorm_bn or pdnorm_ln:
            raise NotImplementedError("PDNorm is disabled in every SceneSplat lang config and is out of scope here")
        self._before_layers(enc_channels)  # (sub-classes register parameters that precede the layers in the reference)
        bn_layer = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        ln_layer = nn.LayerNorm
        act_layer = nn.GELU

        self.embedding = Embedding(in_channels=in_channels, embed_channels=enc_channels[0], norm_layer=bn_layer,
                                   act_layer=act_layer)
        enc_drop_path = [x.item() for x in torch.linspace(0, drop_path, sum(enc_depths))]
        self.enc = PointSequential()
        for s in range(self.num_stages):
            enc_drop_path_ = enc_drop_path[sum(enc_depths[:s]): sum(enc_depths[: s + 1])]
            enc = PointSequential()
            if s > 0:
                enc.add(SerializedPooling(in_channels=enc_channels[s - 1], out_channels=enc_channels[s],
                                          stride=stride[s - 1], norm_layer=bn_layer, act_layer=act_layer,
                                          reduce=self._pooling_reduce()), name="down")
            for i in range(enc_depths[s]):
                enc.add(Block(channels=enc_channels[s], num_heads=enc_num_head[s], patch_size=enc_patch_size[s],
                              mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=proj_drop, drop_path=enc_drop_path_[i], norm_layer=ln_layer, act_layer=act_layer,
                              pre_norm=pre_norm, order_index=i % len(self.order), cpe_indice_key=f"stage{s}",
                              enable_rpe=enable_rpe, enable_flash=enable_flash, upcast_attention=upcast_attention,
                              upcast_softmax=upcast_softmax), name=f"block{i}")
            if len(enc) != 0:
                self.enc.add(module=enc, name=f"enc{s}")
        if self._has_decoder():
            dec_drop_path = [x.item() for x in torch.linspace(0, drop_path, sum(dec_depths))]
            self.dec = PointSequential()
            dec_channels = list(dec_channels) + [enc_channels[-1]]
            for s in reversed(range(self.num_stages - 1)):
                dec_drop_path_ = dec_drop_path[sum(dec_depths[:s]): sum(dec_depths[: s + 1])]
                dec_drop_path_.reverse()
                dec = PointSequential()
                dec.add(SerializedUnpooling(in_channels=dec_channels[s + 1], skip_channels=enc_channels[s],
                                            out_channels=dec_channels[s], norm_layer=bn_layer, act_layer=act_layer),
                        name="up")
                for i in range(dec_depths[s]):
                    dec.add(Block(channels=dec_channels[s], num_heads=dec_num_head[s], patch_size=dec_patch_size[s],
                                  mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                                  proj_drop=proj_drop, drop_path=dec_drop_path_[i], norm_layer=ln_layer,
                                  act_layer=act_layer, pre_norm=pre_norm, order_index=i % len(self.order),
                                  cpe_indice_key=f"stage{s}", enable_rpe=enable_rpe, enable_flash=enable_flash,
                                  upcast_attention=upcast_attention, upcast_softmax=upcast_softmax), name=f"block{i}")
                self.dec.add(module=dec, name=f"dec{s}")

    # hooks of the SSL variant (PointTransformerV3SimDINO)
    def _before_layers(self, enc_channels):
        pass

    def _pooling_reduce(self):
        return "mean"

    def _has_decoder(self):
        return not self.cls_mode

    def plan_indices(self, point):
        """Builds the whole feature-independent index hierarchy up front: pooled codes / orders / clusters of every
        encoder level, the 3^3 kernel maps and the pair lists of every level.  These are the only steps of a forward
        that need a host sync (coarse voxel counts, pair counts); doing them first, on kernels that take well under
        a millisecond each, leaves the feature path (all the heavy kernels) free of syncs, so the host enqueues
        ahead of the device instead of draining it 7 times per forward.  The CPU RNG is consumed in the reference's
        order (serialization, then one randperm per pooling level)."""
        from .spconv_compat import kernel_map_for
        if not any(isinstance(m, Block) and m.cpe[0].tensor_core_ok() for m in self.modules()):
            return
        stem = self.embedding.stem._modules.get("conv")
        if getattr(stem, "kernel_size", 3) == 5 and not stem.tensor_core_ok():
            # the stem's 5^3 map first (here, in the index phase, instead of lazily in front of the stem conv): the level's
            # 3^3 map is then a row subset of it (ss_kmap_subset) and needs no search of its own
            kernel_map_for(point, 5, want_pairs=False)
        kernel_map_for(point, 3, want_pairs=True)
        level = point
        holder = point  # Dict that receives the plan of the next pooling
        for stage in self.enc.children():
            down = getattr(stage, "down", None)
            if down is None:
                continue
            pooling_depth = down.pooling_depth_for(level.serialized_depth)
            perm, ix = down.pool_indices(level, pooling_depth)
            names = level.serialized_order_names
            child = Dict(grid_coord=ix["grid_coord"], batch=ix["batch"], serialized_code=ix["code"],
                         serialized_order=ix["order"], serialized_depth=level.serialized_depth - pooling_depth,
                         serialized_order_names=tuple(names[i] for i in perm))
            kernel_map_for(child, 3, want_pairs=True)
            plan = dict(pooling_depth=pooling_depth, perm=perm, ix=ix, kmap_cache=child["_kmap_cache"], child=None)
            if holder is point:
                point["_pool_plan"] = plan
            else:
                holder["child"] = plan
            holder = plan
            level = child

    def prepare(self, data_dict):
        """Index phase of a forward: serialization + the whole feature-independent index hierarchy (all the host
        syncs of a forward live here).  May run on a side stream while the feature phase of the previous chunk
        occupies the device (`ChunkPipeline`)."""
        point = Point(data_dict)
        point.serialization(order=self.order, shuffle_orders=self.shuffle_orders)
        point.sparsify()
        self.plan_indices(point)
        return point

    fused_l2_normalize = True  # run(point, l2_normalize=True) exists (LangPretrainer asks for it in eval mode)

    def run(self, point, l2_normalize=False):
        """Feature phase of a forward on a prepared Point: no host sync.  l2_normalize (LangPretrainer, eval): the last
        decoder Block writes F.normalize(feat, p=2, dim=1) instead of feat and marks the Point `_l2_normalized`."""
        point = self.embedding(point)
        point = self.enc(point)
        if not self.cls_mode:
            stages = list(self.dec.children())
            for i, stage in enumerate(stages):
                if l2_normalize and i == len(stages) - 1:
                    mods = list(stage.children())
                    for j, m in enumerate(mods):
                        if j == len(mods) - 1 and isinstance(m, Block):
                            point["_emit_l2_normalized"] = True
                        point = _apply(m, point)
                else:
                    point = stage(point)
        return point

    def forward(self, data_dict):
        if torch.is_grad_enabled() and self.training:
            # training: autograd around the package's kernels (scenesplat_b200/training.py)
            from . import training
            return Point(Dict(feat=training.forward_train(self, data_dict)))
        return self.run(self.prepare(data_dict))


@MODELS.register_module("PT-v3m1-simdino")
class PointTransformerV3SimDINO(PointTransformerV3):
    """The self-supervised variant of PT-v3m1 (pointcept/models/point_transformer_v3_ssl/point_transformer_v3m1_ssl.py:
    532-790, registered as "PT-v3m1-simdino"): same encoder / decoder modules and `state_dict` keys, plus
      * `do_mask`: a learnable `mask_token` [1, enc_channels[0]] that replaces the embedded features of masked points
        (ref :585-592, :768-772); the decoder exists only when do_mask is set (ref :676);
      * `pooling_reduce` (default "max") for every SerializedPooling (ref :359, :644);
      * `forward(data_dict, mask=None, return_dec=False)` -> (Point(feat, offset) of the encoder output, decoder Point or
        None) (ref :759-790).
    Inference path (the SSL losses / teacher-student training loop of the reference are outside the hot path)."""

    def __init__(self, *args, do_mask=False, pooling_reduce="max", **kwargs):
        self.do_mask, self.pooling_reduce = do_mask, pooling_reduce
        super().__init__(*args, **kwargs)

    def __setattr__(self, name, value):
        # do_mask / pooling_reduce are set before nn.Module.__init__ ran (the hooks below need them during construction)
        if name in ("do_mask", "pooling_reduce") and "_parameters" not in self.__dict__:
            object.__setattr__(self, name, value)
        else:
            super().__setattr__(name, value)

    def _before_layers(self, enc_channels):
        if self.do_mask:  # created before the layers, like the reference: same RNG stream, same state_dict order
            mask_token = torch.nn.Parameter(torch.zeros(1, enc_channels[0]))
            torch.nn.init.trunc_normal_(mask_token, std=0.02)
            self.mask_token = mask_token

    def _pooling_reduce(self):
        return self.pooling_reduce

    def _has_decoder(self):
        return bool(self.do_mask)

    def forward(self, data_dict, mask=None, return_dec=False):
        if torch.is_grad_enabled() and self.training:
            raise NotImplementedError("PT-v3m1-simdino: only the inference path is built (SURVEY.md 8f, row 4)")
        point = self.prepare(data_dict)
        point = self.embedding(point)
        if mask is not None:
            # ref :768-772: in place on the tensor that point.feat AND sparse_conv_feat.features share
            feat = point.feat
            feat[mask] = self.mask_token.to(device=feat.device, dtype=feat.dtype)
            point.feat = feat
        point = self.enc(point)
        point_enc = Point(Dict(feat=point.feat, offset=point.offset))
        point_dec = self.dec(point) if return_dec else None
        return point_enc, point_dec
