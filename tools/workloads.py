"""Secondary workloads of bench.py (BASELINE.json configs 1, 3, 4, 5) and the same-box library bars that the
survey names as "the kernels to beat" (SURVEY.md sections 2.3 / 8d).  Everything here is measurement code: it drives
the package's public API and times it; nothing here is on the product path.

  cpu_index_config1      config 1: GridSample(0.02) + 4-order serialization + 3 pooling index levels + a segment mean on a
                         ~100 k-Gaussian chunk, CPU (oracle/ = the numpy restatement of the reference's torch / numpy code)
  gpu_index_config1      the same stages through the package's kernels, per-stage CUDA-event times and HBM GB/s
  torch_index_config1    the same stages as the reference writes them in torch (argsort / unique / scatter_), on the GPU
  attention_library_bar  flash-attn 2.8.3 varlen + the two row gathers (ptv3:188,208-216) against the own kernel
  train_step             config 4: lang-pretraining step, 8 chunks per global batch split over the ranks, DDP
  zero_shot_scene        config 3: 1.5 M-Gaussian scene through scenesplat_b200.inference.zero_shot_scene
  sweep                  config 5: 0.5 - 4 M Gaussian scenes, chunked by the 6 m / 3 m rule, LPT-sharded over the ranks
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _ev():
    return torch.cuda.Event(enable_timing=True)


def _time_gpu(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = _ev(), _ev()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


# ------------------------------------------------------------------------------------------------- config 1
CONFIG1_RAW = 100000


def _config1_chunk():
    from scenesplat_b200 import synthetic
    return synthetic.chunk(CONFIG1_RAW, L=6.0, H=3.0, seed=0)


def config1_bytes(n_raw, sizes, c=64):
    """Algorithmic bytes: GridSample 20 N_raw + 136 M; serialization 128 B per voxel; pooling level
    N (28 + C e) + M (148 + 4 C) with e = 4 (fp32 features) for the level that also reduces a [N, C] tensor, index part
    only (N 28 + M 148) for the others."""
    m0 = sizes[0]
    b = dict(gridsample=20.0 * n_raw + 136.0 * m0, serialization=128.0 * m0)
    pool = 0.0
    for i in range(len(sizes) - 1):
        n, m = sizes[i], sizes[i + 1]
        pool += n * 28.0 + m * 148.0
        if i == 0:
            pool += n * c * 4.0 + m * c * 4.0
    b["pooling"] = pool
    b["total"] = sum(b.values())
    return b


def cpu_index_config1(threads=None, reps=3):
    """Reference-style CPU path, best of `reps` after one warm-up (SURVEY.md 8d "CPU reference timing")."""
    from oracle import gridsample as ogs
    from oracle import pooling as opool
    from oracle import serialization as oser
    torch.set_num_threads(threads or os.cpu_count() or 1)
    d = _config1_chunk()
    rng = np.random.default_rng(0)

    def run():
        t = {}
        t0 = time.perf_counter()
        res = ogs.grid_sample_train(d["coord"], 0.02)
        g = res["grid_coord"]
        t["gridsample"] = time.perf_counter() - t0
        n = g.shape[0]
        t0 = time.perf_counter()
        batch = oser.offset2batch(np.array([n]))
        code, order, inv, depth = oser.serialization(g, batch, 1, oser.ORDERS)
        t["serialization"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        sizes = [n]
        feat = rng.standard_normal((n, 64)).astype(np.float32)
        c = code
        for lvl in range(3):
            ix = opool.pool_index(c, 1)
            if lvl == 0:
                opool.segment_csr(feat, ix["indices"], ix["idx_ptr"], "mean")
            c = ix["code"]
            sizes.append(c.shape[1])
        t["pooling"] = time.perf_counter() - t0
        return t, sizes

    run()
    best, sizes = None, None
    for _ in range(reps):
        t, sizes = run()
        if best is None or sum(t.values()) < sum(best.values()):
            best = t
    total = sum(best.values())
    by = config1_bytes(CONFIG1_RAW, sizes)
    return dict(ms={k: 1e3 * v for k, v in best.items()}, total_ms=1e3 * total, gbs=by["total"] / total / 1e9,
                gaussians_per_s=CONFIG1_RAW / total, cores=torch.get_num_threads(), kind="port",
                sample=f"oracle/ numpy restatement of GridSample + Point.serialization + SerializedPooling index x3 + "
                       f"segment mean [N,64] on a {CONFIG1_RAW}-Gaussian chunk ({sizes[0]} voxels), best of {reps}",
                sizes=sizes)


def gpu_index_config1(dev, reps=5):
    """The same stages through the package's kernels (device-resident input), CUDA-event time per stage."""
    import scenesplat_b200 as S
    from scenesplat_b200 import ops
    from scenesplat_b200.structure import Point
    d = _config1_chunk()
    coord = torch.from_numpy(d["coord"]).to(dev)
    attrs = {k: torch.from_numpy(d[k]).to(dev) for k in ("color", "opacity", "quat", "scale")}
    gs = S.GridSample(grid_size=0.02, hash_type="fnv", mode="train", keys=("coord", "color", "opacity", "quat", "scale"),
                      return_grid_coord=True, return_inverse=True, device=dev)
    state = {}

    def s_grid():
        np.random.seed(0)
        state["out"] = gs(dict(coord=coord, **attrs))

    def s_ser():
        o = state["out"]
        n = o["coord"].shape[0]
        p = Point(dict(coord=o["coord"], grid_coord=o["grid_coord"], feat=o["coord"], offset=torch.tensor([n], device=dev)))
        p.serialization(order=("z", "z-trans", "hilbert", "hilbert-trans"), shuffle_orders=False)
        state["p"] = p

    def s_pool():
        p = state["p"]
        n = p.coord.shape[0]
        feat = state.setdefault("feat", torch.randn(n, 64, device=dev))
        code, order, gc, batch = p.serialized_code, p.serialized_order, p.grid_coord, p.batch
        sizes = [n]
        for lvl in range(3):
            ix = ops.pool_index(code, order, gc, batch, 1, [0, 1, 2, 3])
            if lvl == 0:
                ops.segment_reduce(feat, order[0].contiguous(), ix["seg_start"], "mean")
            code, order, gc, batch = ix["code"], ix["order"], ix["grid_coord"], ix["batch"]
            sizes.append(ix["m"])
        state["sizes"] = sizes

    ms = {}
    for name, fn in (("gridsample", s_grid), ("serialization", s_ser), ("pooling", s_pool)):
        ms[name] = _time_gpu(fn, reps)
    by = config1_bytes(CONFIG1_RAW, state["sizes"])
    total = sum(ms.values())
    return dict(ms=ms, total_ms=total, gbs=by["total"] / (total * 1e-3) / 1e9,
                gbs_by_stage={k: by[k] / (ms[k] * 1e-3) / 1e9 for k in ms}, bytes=by, sizes=state["sizes"],
                note="includes the host syncs of the public API (voxel count, depth, coarse counts)")


def torch_index_config1(dev, reps=5):
    """Same-box library bar: the reference's own torch op sequence for serialization order/inverse and the pooling
    index build (structure.py:85-92, ptv3:384-412: argsort / unique / sort / scatter_), on the GPU.  The code VALUES
    come from the package's encoder (the reference's hilbert encoder is hundreds of tiny kernels; that is reported
    separately as `cpu`), so this bar isolates sort / unique / scatter."""
    from scenesplat_b200 import ops
    from scenesplat_b200.structure import Point
    import scenesplat_b200 as S
    d = _config1_chunk()
    gs = S.GridSample(grid_size=0.02, hash_type="fnv", mode="train", keys=("coord",), return_grid_coord=True, device=dev)
    np.random.seed(0)
    o = gs(dict(coord=torch.from_numpy(d["coord"]).to(dev)))
    n = o["coord"].shape[0]
    p = Point(dict(coord=o["coord"], grid_coord=o["grid_coord"], feat=o["coord"], offset=torch.tensor([n], device=dev)))
    p.serialization(order=("z", "z-trans", "hilbert", "hilbert-trans"), shuffle_orders=False)
    code0 = p.serialized_code.clone()
    feat = torch.randn(n, 64, device=dev)

    def order_inverse(code):
        order = torch.argsort(code)
        inverse = torch.zeros_like(order).scatter_(
            dim=1, index=order, src=torch.arange(0, code.shape[1], device=order.device).repeat(code.shape[0], 1))
        return order, inverse

    def s_ser():
        order_inverse(code0)

    def s_pool():
        code = code0
        f = feat
        for lvl in range(3):
            code = code >> 3
            code_, cluster, counts = torch.unique(code[0], sorted=True, return_inverse=True, return_counts=True)
            _, indices = torch.sort(cluster)
            idx_ptr = torch.cat([counts.new_zeros(1), torch.cumsum(counts, dim=0)])
            head = indices[idx_ptr[:-1]]
            code = code[:, head]
            order_inverse(code)
            if lvl == 0:
                torch.segment_reduce(f[indices], "mean", lengths=counts)

    ms = dict(serialization=_time_gpu(s_ser, reps), pooling=_time_gpu(s_pool, reps))
    return dict(ms=ms, note="torch.argsort / unique / sort / scatter_ / segment_reduce as the reference writes them "
                            "(structure.py:85-92, ptv3:384-412), codes precomputed")


# ------------------------------------------------------------------------------------------------- attention bar
def attention_library_bar(dev, shapes=((299277, 16, 48), (119000, 16, 32), (299277, 2, 16)), K=1024, reps=5):
    """flash-attn 2.8.3 `flash_attn_varlen_qkvpacked_func` + the `qkv[order]` and `feat[inverse]` gathers the
    reference wraps around it (ptv3:188,208-216), against the own kernel (gathers fused) on the same tensors."""
    from scenesplat_b200 import ops
    out = []
    try:
        from flash_attn import flash_attn_varlen_qkvpacked_func
    except Exception as e:  # pragma: no cover
        flash_attn_varlen_qkvpacked_func = None
        err = repr(e)[:200]
    for n, H, d in shapes:
        C = H * d
        torch.manual_seed(0)
        qkv = torch.randn(n, 3 * C, device=dev).bfloat16()
        order = torch.randperm(n, device=dev)
        inverse = torch.empty_like(order)
        inverse[order] = torch.arange(n, device=dev)
        offset = torch.tensor([n], device=dev)
        table = ops.patch_table(offset, K, n)
        scale = d ** -0.5
        own = _time_gpu(lambda: ops.patch_attention(qkv, order, table, K, H, scale, impl="tc"), reps)
        rec = dict(n=n, heads=H, head_dim=d, own_ms=own, own_tflops=4.0 * K * C * n / own / 1e9)
        if flash_attn_varlen_qkvpacked_func is not None:
            # reference padding rule (ptv3:114-170) for one batch item: last patch = window [n - K, n)
            npatch = (n + K - 1) // K
            pad = torch.arange(npatch * K, device=dev)
            if n % K:
                pad[(npatch - 1) * K:] = torch.arange(n - K, n, device=dev)
            unpad = torch.arange(n, device=dev)
            if n % K:
                r = n - (npatch - 1) * K
                unpad[(npatch - 1) * K:] = torch.arange(npatch * K - r, npatch * K, device=dev)
            cu = torch.arange(0, npatch * K + 1, K, device=dev, dtype=torch.int32)
            o_pad, i_unpad = order[pad], unpad[inverse]

            def lib():
                x = qkv[o_pad]
                y = flash_attn_varlen_qkvpacked_func(x.reshape(-1, 3, H, d), cu, max_seqlen=K, dropout_p=0.0,
                                                     softmax_scale=scale).reshape(-1, C)
                return y[i_unpad]

            def lib_core():
                return flash_attn_varlen_qkvpacked_func(xg.reshape(-1, 3, H, d), cu, max_seqlen=K, dropout_p=0.0,
                                                        softmax_scale=scale)
            try:
                xg = qkv[o_pad]
                got = ops.patch_attention(qkv, order, table, K, H, scale, impl="tc")
                want = lib()
                rec["max_abs_diff_vs_flash"] = (got.float() - want.float()).abs().max().item()
                rec["flash_ms"] = _time_gpu(lib, reps)
                rec["flash_core_ms"] = _time_gpu(lib_core, reps)
                rec["speedup_vs_flash"] = rec["flash_ms"] / own
            except Exception as e:  # pragma: no cover
                rec["flash_error"] = repr(e)[:200]
        else:
            rec["flash_error"] = err
        out.append(rec)
    return out


# ------------------------------------------------------------------------------------------------- config 4
def train_step(dev, rank, world, backbone_cfg, total_chunks=8, n_raw=180000, steps=3, warmup=2, max_chunks_per_pass=4):
    """Lang-pretraining step (pointcept/engines/train.py:196-232): `total_chunks` chunks per GLOBAL batch split over
    the ranks (strong scaling), cosine + L2 + aggregated contrastive loss, AdamW, DDP all-reduce of the 91.7 M
    gradients.  A rank with more than `max_chunks_per_pass` chunks accumulates gradients over micro-batches (the
    activations of 8 chunks do not fit 180 GB).  Returns ms/step (max over ranks), Gaussians/s, and the exposed
    all-reduce time = step with DDP sync - step under no_sync()."""
    import torch.distributed as dist
    import scenesplat_b200 as S
    from scenesplat_b200 import synthetic
    per_rank = max(1, total_chunks // world)
    torch.manual_seed(0)
    model = S.LangPretrainer(backbone=dict(backbone_cfg), criteria=[
        dict(type="CosineSimilarity", reduction="mean", loss_weight=1.0),
        dict(type="L2Loss", reduction="mean", loss_weight=1.0),
        dict(type="AggregatedContrastiveLoss", temperature=0.2, reduction="mean", loss_weight=0.02,
             schedule="last_75")]).to(dev).train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index]) if world > 1 else model
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.05)
    gs = S.GridSample(grid_size=0.02, hash_type="fnv", mode="train", keys=("coord", "color", "opacity", "quat", "scale"),
                      return_grid_coord=True, device=dev)
    micro = []
    for m0 in range(0, per_rank, max_chunks_per_pass):
        parts, sizes = [], []
        for c in range(m0, min(per_rank, m0 + max_chunks_per_pass)):
            d = synthetic.chunk(n_raw, seed=rank * 100 + c)
            np.random.seed(c)
            out = gs({k: torch.from_numpy(v) for k, v in d.items() if k in ("coord", "color", "opacity", "quat", "scale")})
            parts.append(out)
            sizes.append(out["coord"].shape[0])
        n = sum(sizes)
        g = torch.Generator().manual_seed(rank * 10 + m0)
        micro.append(dict(
            coord=torch.cat([p["coord"] for p in parts]), grid_coord=torch.cat([p["grid_coord"] for p in parts]),
            feat=torch.cat([torch.cat([p["color"], p["opacity"], p["quat"], p["scale"]], 1) for p in parts]).contiguous(),
            offset=torch.tensor(np.cumsum(sizes), device=dev),
            lang_feat=torch.nn.functional.normalize(torch.randn(n, 768, generator=g), dim=1).to(dev),
            valid_feat_mask=(torch.rand(n, generator=g) < 0.8).to(dev),
            segment=torch.randint(-1, 200, (n,), generator=g).to(dev), epoch_progress=0.9))
    n_rank = sum(int(b["coord"].shape[0]) for b in micro)

    def step(sync=True):
        opt.zero_grad(set_to_none=True)
        loss = None
        for i, b in enumerate(micro):
            last = i == len(micro) - 1
            if world > 1 and (not sync or not last):
                with net.no_sync():
                    loss = net(dict(b))["loss"] / len(micro)
                    loss.backward()
            else:
                loss = net(dict(b))["loss"] / len(micro)
                loss.backward()
        opt.step()
        return loss

    def timed(sync):
        for _ in range(warmup):
            step(sync)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = _ev(), _ev()
        a.record()
        for _ in range(steps):
            loss = step(sync)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, float(loss.detach())

    ms, loss = timed(True)
    tot = torch.tensor([float(n_rank)], device=dev)
    if world > 1:
        dist.all_reduce(tot)
    out = dict(ms_per_step=ms, gaussians_per_s=float(tot.item()) / (ms * 1e-3), chunks_per_global_batch=per_rank * world,
               chunks_per_rank=per_rank, micro_batches_per_rank=len(micro), voxels_global=int(tot.item()), loss=loss,
               scaling="strong", grad_allreduce_bytes=4 * sum(p.numel() for p in model.parameters()),
               peak_memory_gib=torch.cuda.max_memory_allocated(dev) / 2 ** 30)
    if world > 1:
        ms_nosync, _ = timed(False)
        out["ms_per_step_no_allreduce"] = ms_nosync
        out["exposed_allreduce_ms"] = ms - ms_nosync
    del opt, net, model, micro
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------- config 3
def zero_shot_scene(dev, model, text, n_raw=1500000, L=14.0, reps=1):
    from scenesplat_b200 import inference, synthetic
    d = synthetic.chunk(n_raw, L=L, H=3.0, seed=0)
    scene = {k: torch.from_numpy(v).to(dev) for k, v in d.items() if k in inference.ATTR_KEYS}
    K = text.shape[0]
    gt = torch.from_numpy(((d["coord"][:, 0] * 1.3 + d["coord"][:, 1] * 0.7).astype(np.int64)) % K).to(dev)
    inference.zero_shot_scene(model, scene, text, gt)  # warm-up
    best = None
    for _ in range(reps):
        t = {}
        inference.zero_shot_scene(model, scene, text, gt, timings=t)
        tot = sum(v for k, v in t.items() if k not in ("fragments", "voxels_forwarded"))
        if best is None or tot < best[0]:
            best = (tot, t)
    tot, t = best
    return dict(gaussians=n_raw, fragments=t["fragments"], voxels_forwarded=t["voxels_forwarded"], total_ms=1e3 * tot,
                scene_gaussians_per_s=n_raw / tot,
                stage_ms={k: 1e3 * v for k, v in t.items() if k not in ("fragments", "voxels_forwarded")})


# ------------------------------------------------------------------------------------------------- config 5
def sweep(dev, rank, world, model, text, sizes=(500000, 1000000, 2000000, 4000000), density=12000.0, cpu_rate=None):
    """Scenes of 0.5 - 4 M Gaussians (floor area grows with the count: `density` Gaussians per m^2 of floor, the
    synthetic room's ~10 k/m^2), chunked by the 6 m x 6 m / 3 m rule, GridSample(0.02) per chunk, LPT-sharded over the
    ranks, labels all-gathered.  Per scene: chunks, voxels, per-rank device ms, imbalance, Gaussians/s (scene
    Gaussians / slowest rank).  `cpu_rate` (Gaussians/s of the CPU baseline) gives the host-CPU column."""
    import torch.distributed as dist
    import scenesplat_b200 as S
    from scenesplat_b200 import inference, synthetic
    gs = S.GridSample(grid_size=0.02, hash_type="fnv", mode="train", keys=inference.ATTR_KEYS, return_grid_coord=True,
                      device=dev)
    rows = []
    for n_raw in sizes:
        L = max(6.5, float(np.sqrt(n_raw / density)))
        d = synthetic.chunk(n_raw, L=L, H=3.0, seed=7)  # every rank builds the same scene (seeded)
        scene = {k: torch.from_numpy(v).to(dev) for k, v in d.items() if k in inference.ATTR_KEYS}
        idxs = inference.scene_chunks(scene["coord"])
        chunks = []
        for ci, idx in enumerate(idxs):
            np.random.seed(ci)
            o = gs({k: v[idx] for k, v in scene.items()})
            n = o["coord"].shape[0]
            chunks.append(dict(coord=o["coord"], grid_coord=o["grid_coord"], feat=inference.feat_of(o),
                               offset=torch.tensor([n], device=dev)))
        if not chunks:
            continue
        del scene
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        _, st_cold = inference.sharded_chunk_labels(model, chunks, text, rank, world, "lpt", gather=False)
        if world > 1:
            dist.barrier()
        labels, st = inference.sharded_chunk_labels(model, chunks, text, rank, world, "lpt")
        ms_all = [st["ms"]]
        if world > 1:
            t = torch.zeros(world, device=dev)
            t[rank] = st["ms"]
            dist.all_reduce(t)
            ms_all = t.tolist()
        vox = [int(c["coord"].shape[0]) for c in chunks]
        per_rank_vox = [sum(vox[i] for i in tb) for tb in st["table"]]
        slow = max(ms_all)
        row = dict(scene_gaussians=n_raw, room_edge_m=round(L, 1), chunks=len(chunks), chunk_voxels_min=min(vox),
                   chunk_voxels_max=max(vox), voxels_total=sum(vox), per_rank_ms=[round(m, 2) for m in ms_all],
                   per_rank_voxels=per_rank_vox, imbalance=max(per_rank_vox) / (sum(per_rank_vox) / world),
                   voxels_per_s=sum(vox) / (slow * 1e-3), first_pass_ms_rank0=round(st_cold["ms"], 2),
                   all_gather_ms=st.get("all_gather_ms"),
                   labels_complete=all(l is not None for l in labels))
        if cpu_rate:
            row["host_cpu_s_estimate"] = sum(vox) / cpu_rate
            row["speedup_vs_host_cpu"] = (sum(vox) / (slow * 1e-3)) / cpu_rate
        rows.append(row)
        del chunks, labels
    return rows
