"""Tensor-level wrappers over the C-ABI kernels: allocate outputs with torch, pass raw pointers.

Every function here runs on the current CUDA stream and never falls back to torch ops for the
computation itself (the CUDA library is the only implementation).
"""
from __future__ import annotations

import weakref

import torch

from . import _lib as L

ORDER_IDS = {"z": 0, "z-trans": 1, "hilbert": 2, "hilbert-trans": 3}
_BF16 = torch.bfloat16


def _isbf(t):
    if t.dtype == torch.bfloat16:
        return 1
    if t.dtype == torch.float32:
        return 0
    raise L.CudaKernelError(f"unsupported dtype {t.dtype} (float32 or bfloat16)")


def _i64(t):
    return t if t.dtype == torch.int64 else t.long()


# ------------------------------------------------------------------------------------------- serialization
def coord_depth(grid_coord: torch.Tensor) -> int:
    """depth = bit_length(max grid_coord) (structure.py:66).  One 8-byte D2H read."""
    g = grid_coord.contiguous()
    out = torch.empty(1, dtype=torch.int64, device=g.device)
    L.call("ss_coord_max", L.ptr(g), int(g.dtype == torch.int32), g.shape[0], L.ptr(out), L.stream())
    return max(int(out.item()).bit_length(), 1)


def serialize(grid_coord, offset, depth: int, orders, want_batch=True):
    """-> (batch [n] or None, code [k,n], order [k,n], inverse [k,n]) int64."""
    g = grid_coord.contiguous()
    if g.dtype not in (torch.int64, torch.int32):
        g = g.long()
    offset = _i64(offset).contiguous()
    n, k = g.shape[0], len(orders)
    dev = g.device
    ids = L.int_array([ORDER_IDS[o] if isinstance(o, str) else int(o) for o in orders])
    code = torch.empty((k, n), dtype=torch.int64, device=dev)
    order = torch.empty((k, n), dtype=torch.int64, device=dev)
    inverse = torch.empty((k, n), dtype=torch.int64, device=dev)
    batch = torch.empty(n, dtype=torch.int64, device=dev) if want_batch else None
    nb = offset.numel()
    wsb = L.load().ss_serialize_workspace_bytes(n, k, depth, nb)
    ws = L.workspace(wsb, dev)
    L.call("ss_serialize", L.ptr(g), int(g.dtype == torch.int32), L.ptr(offset), nb, n, depth, k, ids, L.ptr(batch),
           L.ptr(code), L.ptr(order), L.ptr(inverse), L.ptr(ws), ws.numel(), L.stream(),
           meta=dict(bytes=n * ((12 if g.dtype == torch.int32 else 24) + 8 + 24 * k)))
    return batch, code, order, inverse


# ------------------------------------------------------------------------------------------- GridSample
def gridsample_index(coord, grid_size: float, hash_type: str = "fnv"):
    """-> dict(idx_sort, inverse, start, m (python int; one D2H read), m_dev, min_coord)."""
    coord = coord.contiguous().float()
    n, dev = coord.shape[0], coord.device
    idx_sort = torch.empty(n, dtype=torch.int64, device=dev)
    inverse = torch.empty(n, dtype=torch.int64, device=dev)
    start = torch.empty(n + 1, dtype=torch.int64, device=dev)
    m_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    mn = torch.zeros(3, dtype=torch.int64, device=dev)
    ws = L.workspace(L.load().ss_gridsample_workspace_bytes(n), dev)
    L.call("ss_gridsample_index", L.ptr(coord), n, float(grid_size), 0 if hash_type == "fnv" else 1, L.ptr(idx_sort),
           L.ptr(inverse), L.ptr(start), L.ptr(m_dev), L.ptr(mn), L.ptr(ws), ws.numel(), L.stream())
    m = int(m_dev.item())
    return dict(coord=coord, idx_sort=idx_sort, inverse=inverse, start=start, m=m, m_dev=m_dev, min_coord=mn,
                grid_size=float(grid_size))


def gridsample_select(ix, rnd=None, frag: int = 0, want_grid_coord=True, want_count=False):
    coord = ix["coord"]
    n, m, dev = coord.shape[0], ix["m"], coord.device
    idx_unique = torch.empty(m, dtype=torch.int64, device=dev)
    gc = torch.empty((m, 3), dtype=torch.int64, device=dev) if want_grid_coord else None
    cnt = torch.empty(m, dtype=torch.int64, device=dev) if want_count else None
    if rnd is not None:
        rnd = _i64(rnd).contiguous()
    L.call("ss_gridsample_select", L.ptr(coord), n, ix["grid_size"], L.ptr(ix["min_coord"]), L.ptr(ix["idx_sort"]),
           L.ptr(ix["start"]), L.ptr(ix["m_dev"]), L.ptr(rnd), int(frag), L.ptr(idx_unique), L.ptr(gc), L.ptr(cnt),
           L.stream())
    return idx_unique, gc, cnt


def gather_rows(src, idx):
    src = src.contiguous()
    m = idx.shape[0]
    out = torch.empty((m,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    row_bytes = src.element_size() * (src[0].numel() if src.dim() > 1 else 1)
    L.call("ss_gather_rows", L.ptr(src), row_bytes, L.ptr(idx), None, m, L.ptr(out), L.stream())
    return out


def sphere_crop_order(coord, center):
    """order [n] int64: indices by ascending squared distance to `center` (3 floats on the host), numpy fp32 arithmetic,
    ties by ascending index (transform.py:1493-1499)."""
    coord = coord.contiguous()
    n = coord.shape[0]
    order = torch.empty(n, dtype=torch.int64, device=coord.device)
    dist = torch.empty(n, dtype=torch.int64, device=coord.device)
    ws = L.workspace(L.load().ss_sphere_crop_workspace_bytes(n), coord.device)
    L.call("ss_sphere_crop_order", L.ptr(coord), n, L.float_array([float(c) for c in center]), L.ptr(order), L.ptr(dist),
           L.ptr(ws), ws.numel(), L.stream())
    return order


# ------------------------------------------------------------------------------------------- pooling
_POOL_META = dict(bytes=0.0)


def pool_index(code, order, grid_coord, batch, pooling_depth: int, src_row):
    """-> dict(cluster, seg_start, head, m, code, order, inverse, grid_coord, batch) (children sliced to m)."""
    k, n = code.shape
    dev = code.device
    cluster = torch.empty(n, dtype=torch.int64, device=dev)
    seg_start = torch.empty(n + 1, dtype=torch.int64, device=dev)
    head = torch.empty(n, dtype=torch.int64, device=dev)
    m_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    ccode = torch.empty((k, n), dtype=torch.int64, device=dev)
    corder = torch.empty((k, n), dtype=torch.int64, device=dev)
    cinv = torch.empty((k, n), dtype=torch.int64, device=dev)
    gc = _i64(grid_coord).contiguous() if grid_coord is not None else None
    cgc = torch.empty((n, 3), dtype=torch.int64, device=dev) if gc is not None else None
    cb = torch.empty(n, dtype=torch.int64, device=dev) if batch is not None else None
    ws = L.workspace(L.load().ss_pool_workspace_bytes(n), dev)
    code_c, order_c = code.contiguous(), order.contiguous()  # bound to names: they must outlive the launch
    batch_c = batch.contiguous() if batch is not None else None
    L.call("ss_pool_index", L.ptr(code_c), L.ptr(order_c), L.ptr(gc),
           L.ptr(batch_c), n, k, pooling_depth, L.int_array(src_row), n,
           L.ptr(cluster), L.ptr(seg_start), L.ptr(head), L.ptr(m_dev), L.ptr(ccode), L.ptr(corder), L.ptr(cinv),
           L.ptr(cgc), L.ptr(cb), L.ptr(ws), ws.numel(), L.stream(), meta=_POOL_META)
    m = int(m_dev.item())
    if L.PROFILE is not None and L.PROFILE.get("ss_pool_index"):  # algorithmic bytes need m: N*28 + M*148
        rec = L.PROFILE["ss_pool_index"][-1]
        L.PROFILE["ss_pool_index"][-1] = (rec[0], rec[1], dict(bytes=n * 28.0 + m * 148.0))
    return dict(cluster=cluster, seg_start=seg_start[: m + 1], head=head[:m], m=m,
                code=ccode[:, :m].contiguous(), order=corder[:, :m].contiguous(), inverse=cinv[:, :m].contiguous(),
                grid_coord=cgc[:m] if cgc is not None else None, batch=cb[:m] if cb is not None else None)


_REDUCE = {"sum": 0, "mean": 1, "max": 2, "min": 3}


def segment_reduce(src, order_row, seg_start, reduce="mean", scale=None, shift=None, act=0, out_dtype=None):
    src = src.contiguous()
    m = seg_start.shape[0] - 1
    c = src.shape[1]
    out = torch.empty((m, c), dtype=out_dtype or src.dtype, device=src.device)
    L.call("ss_segment_reduce", L.ptr(src), _isbf(src), L.ptr(order_row), L.ptr(seg_start), None, m, c,
           _REDUCE[reduce], L.ptr(scale), L.ptr(shift), act, L.ptr(out), _isbf(out), L.stream(),
           meta=dict(bytes=src.shape[0] * (8.0 + c * src.element_size()) + m * (8.0 + c * out.element_size())))
    return out


def pool_reduce(src, coord, order_row, seg_start, reduce="mean", scale=None, shift=None, act=0, out_dtype=None):
    """segment_reduce of the projected features and the "mean" of the fp32 coordinates over the same segments, one launch."""
    src, coord = src.contiguous(), coord.contiguous()
    m = seg_start.shape[0] - 1
    c = src.shape[1]
    out = torch.empty((m, c), dtype=out_dtype or src.dtype, device=src.device)
    coord_out = torch.empty((m, 3), dtype=torch.float32, device=src.device)
    L.call("ss_pool_reduce", L.ptr(src), _isbf(src), L.ptr(coord), L.ptr(order_row), L.ptr(seg_start), m, c,
           _REDUCE[reduce], L.ptr(scale), L.ptr(shift), act, L.ptr(out), _isbf(out), L.ptr(coord_out), L.stream(),
           meta=dict(bytes=src.shape[0] * (8.0 + c * src.element_size() + 12.0) + m * (8.0 + c * out.element_size() + 12.0)))
    return out, coord_out


def unpool_gather_add(a, b, cluster, scale_a=None, shift_a=None, scale_b=None, shift_b=None, act=0, out_dtype=None,
                      want_a=False, a_dtype=None):
    a, b = a.contiguous(), b.contiguous()
    n, c = a.shape
    od = out_dtype or a.dtype
    ad = a_dtype or od
    out = torch.empty((n, c), dtype=od, device=a.device)
    out_a = torch.empty((n, c), dtype=ad, device=a.device) if want_a else None
    flags = int(od == _BF16) | (2 if (want_a and ad == _BF16) else 0)
    L.call("ss_unpool_gather_add", L.ptr(a), L.ptr(b), _isbf(a), L.ptr(cluster), n, c, L.ptr(scale_a), L.ptr(shift_a),
           L.ptr(scale_b), L.ptr(shift_b), act, L.ptr(out), L.ptr(out_a), flags, L.stream())
    return out, out_a


# ------------------------------------------------------------------------------------------- submanifold conv
def kmap_build(grid_coord, batch, code_row, order_row, depth: int, order_id: int, k: int):
    """-> (nbr [k^3, n] int32, tap_count [k^3] int64 device)."""
    g = grid_coord.contiguous()
    if g.dtype not in (torch.int64, torch.int32):
        g = g.long()
    n, dev = g.shape[0], g.device
    k3 = k ** 3
    nbr = torch.empty((k3, n), dtype=torch.int32, device=dev)
    cnt = torch.empty(k3, dtype=torch.int64, device=dev)
    ws = L.workspace(L.load().ss_kmap_workspace_bytes(n, k), dev)
    batch_c, code_c, order_c = batch.contiguous(), code_row.contiguous(), order_row.contiguous()
    L.call("ss_kmap_build", L.ptr(g), int(g.dtype == torch.int32), L.ptr(batch_c),
           L.ptr(code_c), L.ptr(order_c), n, depth, order_id, k, L.ptr(nbr), L.ptr(cnt),
           L.ptr(ws), ws.numel(), L.stream())
    return nbr, cnt


def kmap_subset(nbr_from, count_from, k_from: int, k_to: int):
    """(nbr, tap_count) of the centred k_to^3 window out of a k_from^3 kernel map of the same voxels (row copies)."""
    n, dev = nbr_from.shape[1], nbr_from.device
    nbr = torch.empty((k_to ** 3, n), dtype=torch.int32, device=dev)
    cnt = torch.empty(k_to ** 3, dtype=torch.int64, device=dev)
    L.call("ss_kmap_subset", L.ptr(nbr_from), L.ptr(count_from), n, k_from, k_to, L.ptr(nbr), L.ptr(cnt), L.stream())
    return nbr, cnt


def upload(values, dtype, device):
    """Small host table -> device WITHOUT draining the stream: torch.tensor(..., device=cuda) copies from pageable
    memory and synchronises; a pinned staging buffer with non_blocking=True only enqueues the copy (the pinned block
    is recycled by torch's host allocator after the copy has run)."""
    t = torch.tensor(values, dtype=dtype)
    if torch.device(device).type != "cuda":
        return t.to(device)
    return t.pin_memory().to(device, non_blocking=True)


CONV_TILE = 256  # rows per gather-GEMM tile (one CTA pair = 2 x 128, or one CTA with two accumulators)
FUSED_CONV_MIN_C = 512  # from this width on the xCPE conv + gather-sum + LN + residual + LN is ONE launch


def kmap_pairs(nbr, order_row, k: int, tap_count_host, tile: int = None):
    """Pair lists for the gather-GEMM conv.  -> dict(pair_in [p_pad] i32, ypos [k^3, n] i32, tile_tap [p_pad/tile] i32,
    p_pad, pairs, tile, ...).  Every tap's segment is padded to a multiple of `tile` rows.  For the fused conv
    (ss_subm_conv_fused_add_ln; k = 3, 256-row tiles) also: ypos_rank, tile_order (the tiles sorted by the rank along
    `order_row` of their first output, as ss_kmap_pairs reports it), its inverse tile_pos, tile_flags (scratch), order_row."""
    tile = tile or CONV_TILE
    k3, n = nbr.shape
    dev = nbr.device
    base, tile_tap, o = [], [], 0
    for t in range(k3):
        base.append(o)
        nt = (int(tap_count_host[t]) + tile - 1) // tile
        tile_tap += [t] * nt
        o += nt * tile
    p_pad = o
    base_dev = upload(base, torch.int64, dev)
    pair_in = torch.empty(max(p_pad, 1), dtype=torch.int32, device=dev)
    ypos = torch.empty((k3, n), dtype=torch.int32, device=dev)
    fused_ok = k == 3 and tile == 256 and len(tile_tap) > 0
    ypos_rank = torch.empty((n, 32), dtype=torch.int32, device=dev) if fused_ok else None
    first_rank, tile_order, tile_pos = (torch.empty(len(tile_tap), dtype=torch.int32, device=dev) if fused_ok else None
                                        for _ in range(3))
    ws = L.workspace(L.load().ss_kmap_workspace_bytes(n, k), dev)
    order_c = order_row.contiguous()
    L.call("ss_kmap_pairs", L.ptr(nbr), L.ptr(order_c), n, k, L.ptr(base_dev), p_pad, L.ptr(pair_in),
           L.ptr(ypos), L.ptr(ypos_rank), L.ptr(first_rank), L.ptr(tile_order), L.ptr(tile_pos), L.ptr(ws), ws.numel(),
           L.stream())
    return dict(pair_in=pair_in, ypos=ypos, ypos_rank=ypos_rank, tile_tap=upload(tile_tap or [0], torch.int32, dev),
                tile_order=tile_order, tile_pos=tile_pos,
                tile_flags=torch.empty(len(tile_tap) + 2, dtype=torch.int32, device=dev), order_row=order_c,
                p_pad=p_pad, pairs=int(sum(int(c) for c in tap_count_host)), tile=tile, tap_base=list(base),
                tap_count=[int(c) for c in tap_count_host])


def subm_conv_simt(x, nbr, wt, bias=None, scale=None, shift=None, act=0, out_dtype=None):
    """wt: [k^3, cin, cout] fp32."""
    x = x.contiguous()
    k3, n = nbr.shape
    cin, cout = wt.shape[1], wt.shape[2]
    out = torch.empty((n, cout), dtype=out_dtype or x.dtype, device=x.device)
    L.call("ss_subm_conv_simt", L.ptr(x), _isbf(x), L.ptr(nbr), L.ptr(wt), L.ptr(bias), L.ptr(scale), L.ptr(shift), act,
           n, k3, cin, cout, L.ptr(out), _isbf(out), L.stream())
    return out


def _subm_conv_products(x_bf16, pairs, w_bf16, impl="auto"):
    """Stage 1 of the tensor-core conv: prod[r, :] = x[pair_in[r], :] @ w[tap(r)]^T (bf16 [p_pad, cout])."""
    k3, cout, cin = w_bf16.shape
    p_pad = pairs["p_pad"]
    if pairs.get("tile") != CONV_TILE:
        raise L.CudaKernelError("pair lists must be padded to %d-row tiles (ops.kmap_pairs)" % CONV_TILE)
    prod = torch.empty((max(p_pad, 1), cout), dtype=_BF16, device=x_bf16.device)
    pair = cout >= 256 if impl == "auto" else impl == "pair"
    fn = "ss_subm_conv_gemm_pair" if pair else "ss_subm_conv_gemm256"
    x_bf16 = x_bf16.contiguous()
    L.call(fn, L.ptr(x_bf16), L.ptr(pairs["pair_in"]), L.ptr(w_bf16),
           L.ptr(pairs["tile_tap"]), p_pad, k3, cin, cout, L.ptr(prod), L.stream(),
           meta=dict(flops=2.0 * pairs["pairs"] * cin * cout, bytes=2.0 * pairs["pairs"] * (cin + cout)))
    return prod


def subm_conv_gemm(x_bf16, pairs, w_bf16, bias, n: int, out_dtype=torch.bfloat16, impl="auto"):
    """tcgen05 gather-GEMM (+ gather-sum).  w_bf16: [k^3, cout, cin] bf16.  impl: "auto" = CTA pairs for cout >= 256
    (csrc/conv_gemm3.cu), the single-CTA two-accumulator kernel below that (csrc/conv_gemm2.cu); "single" / "pair"
    force one of them (they are bit-identical; tests compare them)."""
    k3, cout, cin = w_bf16.shape
    prod = _subm_conv_products(x_bf16, pairs, w_bf16, impl)
    out = torch.empty((n, cout), dtype=out_dtype, device=x_bf16.device)
    L.call("ss_subm_conv_reduce", L.ptr(prod), L.ptr(pairs["ypos"]), L.ptr(bias), n, k3, cout, L.ptr(out), _isbf(out),
           L.stream(), meta=dict(bytes=2.0 * pairs["pairs"] * cout + n * (4.0 * k3 + out.element_size() * cout)))
    return out


def subm_conv_gemm_add_ln(x_bf16, pairs, w_bf16, bias, res_f32, ln0, ln1, eps=1e-5, inplace=True, impl="auto"):
    """The conv with the Block's next two steps fused into its gather-sum stage: y = res + LN0(conv(x)) (fp32, `res`
    itself when inplace) and LN1(y) (bf16).  The conv output never reaches memory.  -> (y, LN1(y)).
    impl: "auto" = ONE launch for cout >= 512 (the gather-sum runs on reducer warps inside the gather-GEMM and reads the
    products back from L2: csrc/conv_gemm3.cu; measured 2.90 -> 2.62 ms at C = 768, 1.57 -> 1.48 at C = 512, but 0.67 ->
    0.83 at C = 256, where the GEMM is too short to hide the reducers), two launches below that; "fused" / "split" force
    one of the two (bit-identical; tests compare)."""
    k3, cout, cin = w_bf16.shape
    n = res_f32.shape[0]
    if res_f32.dtype != torch.float32 or not res_f32.is_contiguous() or tuple(res_f32.shape) != (n, cout):
        raise L.CudaKernelError("subm_conv_gemm_add_ln: the residual must be a contiguous fp32 [n, cout] tensor")
    (g0, b0), (g1, b1) = ln0, ln1
    fused = impl == "fused" or (impl == "auto" and cout >= FUSED_CONV_MIN_C)
    if fused and cout >= 256 and k3 <= 32 and n > 0 and pairs.get("ypos_rank") is not None:
        if pairs.get("tile") != CONV_TILE:
            raise L.CudaKernelError("pair lists must be padded to %d-row tiles (ops.kmap_pairs)" % CONV_TILE)
        p_pad = pairs["p_pad"]
        prod = torch.empty((max(p_pad, 1), cout), dtype=_BF16, device=x_bf16.device)
        out = res_f32 if inplace else torch.empty_like(res_f32)
        norm = torch.empty((n, cout), dtype=_BF16, device=x_bf16.device)
        x_bf16 = x_bf16.contiguous()
        L.call("ss_subm_conv_fused_add_ln", L.ptr(x_bf16), L.ptr(pairs["pair_in"]), L.ptr(w_bf16), L.ptr(pairs["tile_tap"]),
               L.ptr(pairs["tile_order"]), L.ptr(pairs["tile_pos"]), p_pad, k3, cin, cout, L.ptr(prod),
               L.ptr(pairs["tile_flags"]), L.ptr(pairs["order_row"]), L.ptr(pairs["ypos_rank"]), L.ptr(bias), L.ptr(res_f32), L.ptr(g0), L.ptr(b0), L.ptr(g1),
               L.ptr(b1), float(eps), n, L.ptr(out), L.ptr(norm), L.stream(),
               meta=dict(flops=2.0 * pairs["pairs"] * cin * cout,
                         bytes=2.0 * pairs["pairs"] * cin + n * (4.0 * k3 + 10.0 * cout)))
        return out, norm
    prod = _subm_conv_products(x_bf16, pairs, w_bf16)
    out = res_f32 if inplace else torch.empty_like(res_f32)
    norm = torch.empty((n, cout), dtype=_BF16, device=x_bf16.device)
    L.call("ss_subm_conv_reduce_add_ln", L.ptr(prod), L.ptr(pairs["ypos"]), L.ptr(bias), L.ptr(res_f32), L.ptr(g0), L.ptr(b0),
           L.ptr(g1), L.ptr(b1), float(eps), n, k3, cout, L.ptr(out), L.ptr(norm), L.stream(),
           meta=dict(bytes=2.0 * pairs["pairs"] * cout + n * (4.0 * k3 + 10.0 * cout)))
    return out, norm


# ------------------------------------------------------------------------------------------- attention
def patch_table(offset, patch_size: int, n_total: int):
    offset = _i64(offset).contiguous()
    nb = offset.numel()
    max_patches = n_total // patch_size + nb
    table = torch.empty((max_patches, 4), dtype=torch.int32, device=offset.device)
    L.call("ss_patch_table", L.ptr(offset), nb, patch_size, max_patches, L.ptr(table), None, L.stream())
    return table


def patch_attention(qkv, order_row, table, patch_size: int, heads: int, scale: float, out_dtype=None, impl="auto"):
    qkv = qkv.contiguous()
    n, c3 = qkv.shape
    c = c3 // 3
    d = c // heads
    out = torch.empty((n, c), dtype=out_dtype or qkv.dtype, device=qkv.device)
    tc_ok = qkv.dtype == _BF16 and out.dtype == _BF16 and d in (16, 32, 48) and patch_size <= 1024
    if impl == "tc" and not tc_ok:
        raise L.CudaKernelError("tcgen05 attention needs bf16 in/out, head_dim in {16,32,48}, patch_size <= 1024")
    meta = dict(flops=4.0 * min(patch_size, n) * c * n, bytes=n * (4.0 * c * qkv.element_size() + 8),
                exps=float(min(patch_size, n)) * n * heads)
    if impl == "tc" or (impl == "auto" and tc_ok):
        L.call("ss_patch_attention", L.ptr(qkv), L.ptr(order_row), L.ptr(table), table.shape[0], patch_size, heads, d,
               float(scale), L.ptr(out), L.stream(), meta=meta)
    else:
        L.call("ss_patch_attention_simt", L.ptr(qkv), _isbf(qkv), L.ptr(order_row), L.ptr(table), table.shape[0],
               patch_size, heads, d, float(scale), L.ptr(out), _isbf(out), L.stream(), meta=meta)
    return out


def patch_attention_lse(qkv, order_row, table, patch_size: int, heads: int, scale: float):
    """Training forward (tcgen05 kernel only): -> (out bf16 [n, C], lse2 fp32 [H, n] by sorted position)."""
    qkv = qkv.contiguous()
    n, c3 = qkv.shape
    c = c3 // 3
    d = c // heads
    if not (qkv.dtype == _BF16 and d in (16, 32, 48) and patch_size <= 1024):
        raise L.CudaKernelError("tcgen05 attention needs bf16 in/out, head_dim in {16,32,48}, patch_size <= 1024")
    out = torch.empty((n, c), dtype=_BF16, device=qkv.device)
    lse2 = torch.empty((heads, n), dtype=torch.float32, device=qkv.device)
    L.call("ss_patch_attention_lse", L.ptr(qkv), L.ptr(order_row), L.ptr(table), table.shape[0], patch_size, heads, d,
           float(scale), L.ptr(out), L.ptr(lse2), n, L.stream(),
           meta=dict(flops=4.0 * min(patch_size, n) * c * n, bytes=n * (4.0 * c * 2 + 8),
                     exps=float(min(patch_size, n)) * n * heads))
    return out, lse2


def patch_attention_backward(qkv, out, dout, lse2, order_row, table, patch_size: int, heads: int, scale: float):
    """-> d(qkv) bf16 [n, 3C]: tcgen05 backward of the patch attention incl. the order / inverse gather adjoints."""
    qkv, out, dout = qkv.contiguous(), out.contiguous(), dout.contiguous()
    n, c3 = qkv.shape
    c = c3 // 3
    d = c // heads
    if not (qkv.dtype == _BF16 and out.dtype == _BF16 and dout.dtype == _BF16 and d in (16, 32, 48) and patch_size <= 1024):
        raise L.CudaKernelError("tcgen05 attention backward needs bf16 tensors, head_dim in {16,32,48}, patch_size <= 1024")
    dqkv = torch.empty_like(qkv)
    nbytes = L.load().ss_patch_attention_backward_workspace_bytes(n, heads, d)
    ws = L.workspace(nbytes, qkv.device)
    L.call("ss_patch_attention_backward", L.ptr(qkv), L.ptr(out), L.ptr(dout), L.ptr(lse2), L.ptr(order_row),
           L.ptr(table), table.shape[0], patch_size, heads, d, float(scale), n, L.ptr(dqkv), L.ptr(ws), ws.numel(),
           L.stream(), meta=dict(flops=14.0 * min(patch_size, n) * c * n, bytes=n * (9.0 * c * 2 + 16 * c),
                                 exps=2.0 * min(patch_size, n) * n * heads))
    return dqkv


def linear_act(x_bf16, w_bf16, bias_f32=None, act=0):
    """out = act(x W^T + b) on tcgen05 CTA pairs with the activation in the epilogue (act: 0 none, 1 exact GELU)."""
    x_bf16, w_bf16 = x_bf16.contiguous(), w_bf16.contiguous()
    n, cin = x_bf16.shape
    cout = w_bf16.shape[0]
    out = torch.empty((n, cout), dtype=_BF16, device=x_bf16.device)
    L.call("ss_linear_act_bf16", L.ptr(x_bf16), L.ptr(w_bf16), L.ptr(bias_f32), n, cin, cout, int(act), L.ptr(out), L.stream(),
           meta=dict(flops=2.0 * n * cin * cout, bytes=2.0 * (n * cin + n * cout + cin * cout)))
    return out


def linear_ok(cin: int, cout: int) -> bool:
    """Shapes the CTA-pair GEMM tiles (K in 16-element steps, output columns in 32-column groups)."""
    return cin >= 16 and cin % 16 == 0 and cout >= 32 and cout % 32 == 0


def linear_residual(x_bf16, w_bf16, bias_f32, res_f32, want_bf16=True, inplace=True):
    """res + (x W^T + b) with the add in the GEMM epilogue: -> (fp32 [n, cout] (res itself when inplace), bf16 copy or None)."""
    x_bf16, w_bf16 = x_bf16.contiguous(), w_bf16.contiguous()
    n, cin = x_bf16.shape
    cout = w_bf16.shape[0]
    if res_f32.dtype != torch.float32 or not res_f32.is_contiguous() or tuple(res_f32.shape) != (n, cout):
        raise L.CudaKernelError("linear_residual: the residual must be a contiguous fp32 [n, cout] tensor")
    out = res_f32 if inplace else torch.empty_like(res_f32)
    shadow = torch.empty((n, cout), dtype=_BF16, device=x_bf16.device) if want_bf16 else None
    L.call("ss_linear_residual_bf16", L.ptr(x_bf16), L.ptr(w_bf16), L.ptr(bias_f32), L.ptr(res_f32), n, cin, cout, L.ptr(out),
           L.ptr(shadow), L.stream(),
           meta=dict(flops=2.0 * n * cin * cout, bytes=2.0 * (n * cin + cin * cout) + n * cout * (8.0 + (2.0 if want_bf16 else 0.0))))
    return out, shadow


# ------------------------------------------------------------------------------------------- row-wise fusions
def add_layernorm(res, delta, ln0=None, ln1=None, eps=1e-5, want_res=True, norm_dtype=None, inplace=False):
    """y = res + LN0(delta) (LN0 optional); returns (y fp32 or None, LN1(y) / cast(y) or None)."""
    ref = delta if delta is not None else res
    n, c = ref.shape
    dev = ref.device
    res_out = None
    if want_res:
        res_out = res if (inplace and res is not None) else torch.empty((n, c), dtype=torch.float32, device=dev)
    norm_out = torch.empty((n, c), dtype=norm_dtype, device=dev) if norm_dtype is not None else None
    g0, b0 = ln0 if ln0 is not None else (None, None)
    g1, b1 = ln1 if ln1 is not None else (None, None)
    nbytes = n * c * ((4 if res is not None else 0) + (delta.element_size() if delta is not None else 0)
                      + (4 if res_out is not None else 0) + (norm_out.element_size() if norm_out is not None else 0))
    L.call("ss_add_layernorm", L.ptr(res), L.ptr(delta), _isbf(delta) if delta is not None else 0, L.ptr(g0), L.ptr(b0),
           L.ptr(g1), L.ptr(b1), float(eps), n, c, L.ptr(res_out), L.ptr(norm_out),
           _isbf(norm_out) if norm_out is not None else 0, L.stream(), meta=dict(bytes=float(nbytes)))
    return res_out, norm_out


_BF16_SHADOW = {}  # id(fp32 feature tensor) -> (weak reference to it, its _version, bf16 copy written by the same kernel)


def _set_bf16_shadow(t, shadow):
    key = id(t)  # (tensors compare elementwise: they cannot be WeakKeyDictionary keys)
    _BF16_SHADOW[key] = (weakref.ref(t, lambda _r, k=key: _BF16_SHADOW.pop(k, None)), t._version, shadow)


def bf16_shadow(t):
    """The bf16 copy a fused kernel wrote beside the fp32 tensor `t`, if `t` has not been modified in place since."""
    ent = _BF16_SHADOW.get(id(t))
    return ent[2] if ent is not None and ent[0]() is t and ent[1] == t._version else None


def add_l2_normalize(res_f32, delta, eps=1e-12, want_bf16=True):
    """F.normalize(res + delta, p=2, dim=1, eps) in one pass (fp32 out): the backbone's last residual add fused with
    LangPretrainer's normalisation.  With want_bf16 the kernel also writes the rows rounded to bf16 (what the zero-shot
    head's GEMM reads: lang_head_argmax / lang_head_accumulate pick it up through bf16_shadow, no conversion pass)."""
    n, c = res_f32.shape
    out = torch.empty((n, c), dtype=torch.float32, device=res_f32.device)
    shadow = torch.empty((n, c), dtype=_BF16, device=res_f32.device) if want_bf16 else None
    L.call("ss_add_l2_normalize", L.ptr(res_f32), L.ptr(delta), _isbf(delta), float(eps), n, c, L.ptr(out), L.ptr(shadow),
           L.stream(), meta=dict(bytes=float(n * c * (8 + delta.element_size() + (2 if want_bf16 else 0)))))
    if want_bf16:
        _set_bf16_shadow(out, shadow)
    return out


def subm_conv_wgrad(x_bf16, dy_bf16, pairs, pair_out, k3, rows_per_chunk=8192):
    """-> dw fp32 [k3, cout, cin] = per-tap dY[pair_out]^T X[pair_in] on the tensor cores (split-K over pair chunks)."""
    cin, cout = x_bf16.shape[1], dy_bf16.shape[1]
    key = ("wgrad_chunks", rows_per_chunk)
    if key not in pairs:
        rows = []
        for t, (b0, c) in enumerate(zip(pairs["tap_base"], pairs["tap_count"])):
            for k in range(b0, b0 + c, rows_per_chunk):
                rows.append((t, k, min(k + rows_per_chunk, b0 + c), 0))
        pairs[key] = upload(rows or [(0, 0, 0, 0)], torch.int32, x_bf16.device), len(rows)
    chunks, n_chunks = pairs[key]
    dw = torch.zeros((k3, cout, cin), dtype=torch.float32, device=x_bf16.device)
    x_bf16, dy_bf16 = x_bf16.contiguous(), dy_bf16.contiguous()
    L.call("ss_subm_conv_wgrad", L.ptr(x_bf16), L.ptr(dy_bf16), L.ptr(pairs["pair_in"]),
           L.ptr(pair_out), L.ptr(chunks), n_chunks, k3, cin, cout, L.ptr(dw), L.stream(),
           meta=dict(flops=2.0 * pairs["pairs"] * cin * cout))
    return dw


def gelu_backward(x_bf16, dy_bf16):
    x_bf16, dy_bf16 = x_bf16.contiguous(), dy_bf16.contiguous()
    dx = torch.empty_like(x_bf16)
    L.call("ss_gelu_backward_bf16", L.ptr(x_bf16), L.ptr(dy_bf16), x_bf16.numel(), L.ptr(dx), L.stream())
    return dx


def stem_conv_wgrad(x_f32, dy_f32, nbr, k3):
    """-> dw fp32 [k3, cin, 32] of the stem conv (tiny Cin, Cout = 32), deterministic."""
    x_f32, dy_f32 = x_f32.contiguous(), dy_f32.contiguous()
    n, cin = x_f32.shape
    cout = dy_f32.shape[1]
    dw = torch.empty((k3, cin, cout), dtype=torch.float32, device=x_f32.device)
    ws = L.workspace(L.load().ss_stem_conv_wgrad_workspace_bytes(k3, cin), x_f32.device)
    L.call("ss_stem_conv_wgrad", L.ptr(x_f32), L.ptr(dy_f32), L.ptr(nbr), n, k3, cin, cout, L.ptr(dw), L.ptr(ws), ws.numel(),
           L.stream())
    return dw


def colsum(x_bf16):
    """-> fp32 [C] = x.sum(0) of a bf16 [N, C] matrix (Linear bias gradient), deterministic."""
    x_bf16 = x_bf16.contiguous()
    n, c = x_bf16.shape
    out = torch.empty(c, dtype=torch.float32, device=x_bf16.device)
    ws = L.workspace(L.load().ss_colsum_workspace_bytes(c), x_bf16.device)
    L.call("ss_colsum_bf16", L.ptr(x_bf16), n, c, L.ptr(out), L.ptr(ws), ws.numel(), L.stream())
    return out


def layernorm_backward(x, dy, gamma, eps=1e-5):
    """-> (dx in x's dtype, dgamma fp32 [C], dbeta fp32 [C]) of y = LayerNorm(x; gamma, beta) contracted with dy."""
    x, dy = x.contiguous(), dy.contiguous()
    n, c = x.shape
    dx = torch.empty_like(x)
    dg = torch.zeros(c, dtype=torch.float32, device=x.device)
    db = torch.zeros(c, dtype=torch.float32, device=x.device)
    L.call("ss_layernorm_backward", L.ptr(x), _isbf(x), L.ptr(dy), _isbf(dy), L.ptr(gamma), float(eps), n, c, L.ptr(dx),
           L.ptr(dg), L.ptr(db), L.stream())
    return dx, dg, db


def affine_act(x, scale=None, shift=None, act=0, out_dtype=None):
    x = x.contiguous()
    n, c = x.shape
    out = torch.empty((n, c), dtype=out_dtype or x.dtype, device=x.device)
    L.call("ss_affine_act", L.ptr(x), _isbf(x), L.ptr(scale), L.ptr(shift), act, n, c, L.ptr(out), _isbf(out), L.stream())
    return out


def l2_normalize(x, eps=1e-12, out_dtype=None):
    x = x.contiguous()
    n, c = x.shape
    out = torch.empty((n, c), dtype=out_dtype or x.dtype, device=x.device)
    L.call("ss_l2_normalize", L.ptr(x), _isbf(x), n, c, float(eps), L.ptr(out), _isbf(out), L.stream())
    return out


# ------------------------------------------------------------------------------------------- language head / losses
def _head_tc_ok(feat, text):
    return text.shape[0] <= 256 and feat.shape[1] % 16 == 0 and feat.shape[1] >= 16


def _head_operand(feat, normalize):
    if normalize:
        return l2_normalize(feat, out_dtype=_BF16)
    if feat.dtype == _BF16:
        return feat
    sh = bf16_shadow(feat)
    return sh if sh is not None else feat.to(_BF16)


def lang_head_argmax(feat, text, normalize=False, threshold=0.1, impl="auto"):
    """impl: "simt" = fp32 CUDA-core kernel (exact fp32 logits); "tc" = tcgen05 GEMM with the max/argmax fused
    into the TMEM epilogue (bf16 operands, fp32 accumulate); "auto" = tensor cores when the shape allows."""
    feat = feat.contiguous()
    n, c = feat.shape
    if impl == "tc" or (impl == "auto" and _head_tc_ok(feat, text)):
        fb = _head_operand(feat, normalize)
        tb = text.contiguous().to(_BF16)
        mx = torch.empty(n, dtype=torch.float32, device=feat.device)
        lab = torch.empty(n, dtype=torch.int64, device=feat.device)
        L.call("ss_lang_head_tc", L.ptr(fb), L.ptr(tb), n, c, text.shape[0], float(threshold), 0, None, L.ptr(mx),
               L.ptr(lab), None, L.stream(), meta=dict(flops=2.0 * n * c * text.shape[0]))
        return mx, lab
    text = text.contiguous().float()
    mx = torch.empty(n, dtype=torch.float32, device=feat.device)
    lab = torch.empty(n, dtype=torch.int64, device=feat.device)
    L.call("ss_lang_head", L.ptr(feat), _isbf(feat), L.ptr(text), n, c, text.shape[0], int(normalize), float(threshold),
           0, None, L.ptr(mx), L.ptr(lab), None, L.stream())
    return mx, lab


def lang_head_accumulate(feat, text, probs_accum, idx=None, normalize=False, impl="auto"):
    feat = feat.contiguous()
    n, c = feat.shape
    if impl == "tc" or (impl == "auto" and _head_tc_ok(feat, text)):
        fb = _head_operand(feat, normalize)
        tb = text.contiguous().to(_BF16)  # (bound to a name: the temporary must outlive the call's argument list)
        L.call("ss_lang_head_tc", L.ptr(fb), L.ptr(tb), n, c, text.shape[0], 0.0, 1,
               L.ptr(idx), None, None, L.ptr(probs_accum), L.stream(), meta=dict(flops=2.0 * n * c * text.shape[0]))
        return probs_accum
    text = text.contiguous().float()
    L.call("ss_lang_head", L.ptr(feat), _isbf(feat), L.ptr(text), n, c, text.shape[0], int(normalize), 0.0, 1,
           L.ptr(idx), None, None, L.ptr(probs_accum), L.stream())
    return probs_accum


_TGT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def cos_l2_sums(pred, target, mask):
    """-> double[3] device tensor: sum(1-cos), sum ||p-t||^2, n_valid."""
    pred, target = pred.contiguous(), target.contiguous()
    m8 = mask.contiguous().to(torch.uint8) if mask.dtype != torch.uint8 else mask.contiguous()
    if mask.dtype == torch.bool:
        m8 = mask.contiguous().view(torch.uint8)
    n, c = pred.shape
    acc = torch.empty(3, dtype=torch.float64, device=pred.device)
    L.call("ss_cos_l2_loss", L.ptr(pred), _isbf(pred), L.ptr(target), _TGT[target.dtype], L.ptr(m8), n, c, L.ptr(acc),
           L.stream())
    return acc


def class_half_sums(pred, mask, segment, half, n_classes: int):
    pred = pred.contiguous()
    m8 = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.contiguous().to(torch.uint8)
    n, c = pred.shape
    sums = torch.empty((n_classes * 2, c), dtype=torch.float32, device=pred.device)
    counts = torch.empty(n_classes * 2, dtype=torch.int32, device=pred.device)
    seg64, half64 = _i64(segment).contiguous(), _i64(half).contiguous()  # two live tensors: never the same block
    L.call("ss_class_half_sums", L.ptr(pred), _isbf(pred), L.ptr(m8), L.ptr(seg64),
           L.ptr(half64), n, c, n_classes, L.ptr(sums), L.ptr(counts), L.stream())
    return sums, counts


# ------------------------------------------------------------------------------------------- adjoints (training)
def _mask8(mask):
    return mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.contiguous().to(torch.uint8)


def segment_mean_backward(dout, cluster, seg_start, reduce="mean", out_dtype=None):
    """d src of segment_reduce(src, order0, seg_start, "mean" | "sum"): dsrc[p] = dout[cluster[p]] (/ count)."""
    dout = dout.contiguous()
    n, c = cluster.shape[0], dout.shape[1]
    dsrc = torch.empty((n, c), dtype=out_dtype or dout.dtype, device=dout.device)
    cluster, seg_start = _i64(cluster).contiguous(), _i64(seg_start).contiguous()
    L.call("ss_segment_mean_bwd", L.ptr(dout), _isbf(dout), L.ptr(cluster), L.ptr(seg_start), n, c, _REDUCE[reduce], L.ptr(dsrc),
           _isbf(dsrc), L.stream())
    return dsrc


def unpool_gather_add_backward(dout, order0, seg_start, out_dtype=None):
    """d child of out = a + child[cluster]: the segment sum of dout over each cluster's members."""
    dout = dout.contiguous()
    m, c = seg_start.shape[0] - 1, dout.shape[1]
    dchild = torch.empty((m, c), dtype=out_dtype or dout.dtype, device=dout.device)
    order0, seg_start = _i64(order0).contiguous(), _i64(seg_start).contiguous()
    L.call("ss_unpool_gather_add_bwd", L.ptr(dout), _isbf(dout), L.ptr(order0), L.ptr(seg_start), m, c, L.ptr(dchild),
           _isbf(dchild), L.stream())
    return dchild


def cos_l2_backward(pred, target, mask, acc, grad_out, w_cos, w_l2):
    """d pred (fp32) of w_cos * mean_valid(1 - cos) + w_l2 * mean_valid ||pred - target||^2, times the device scalar grad_out."""
    pred, target, m8 = pred.contiguous(), target.contiguous(), _mask8(mask)
    n, c = pred.shape
    dpred = torch.empty((n, c), dtype=torch.float32, device=pred.device)
    g = grad_out.reshape(1).float().contiguous() if grad_out is not None else None
    L.call("ss_cos_l2_loss_bwd", L.ptr(pred), _isbf(pred), L.ptr(target), _TGT[target.dtype], L.ptr(m8), n, c, L.ptr(acc), L.ptr(g),
           float(w_cos), float(w_l2), L.ptr(dpred), L.stream())
    return dpred


def class_half_sums_backward(dsums, mask, segment, half, n_classes: int):
    dsums = dsums.contiguous().float()
    m8, seg64, half64 = _mask8(mask), _i64(segment).contiguous(), _i64(half).contiguous()
    n, c = seg64.shape[0], dsums.shape[1]
    dpred = torch.empty((n, c), dtype=torch.float32, device=dsums.device)
    L.call("ss_class_half_sums_bwd", L.ptr(dsums), L.ptr(m8), L.ptr(seg64), L.ptr(half64), n, c, n_classes, L.ptr(dpred), L.stream())
    return dpred
