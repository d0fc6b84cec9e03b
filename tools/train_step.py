"""Lang-pretraining training step (BASELINE.json configs[3]): LangPretrainer(PT-v3m1 lang config) with the cosine +
L2 + aggregated contrastive losses on a batch of synthetic chunks, AdamW, DDP gradient all-reduce over NCCL when
launched with torchrun.  Prints Gaussians/s per step (all ranks) and the phase split on rank 0.

  python tools/train_step.py [--chunks 2] [--n-raw 180000] [--steps 5]
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_step.py ...
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import scenesplat_b200 as S
from scenesplat_b200 import synthetic
from bench import LANG_BACKBONE

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", type=int, default=2, help="chunks per rank per step (the reference trains 8 per batch over all ranks)")
ap.add_argument("--n-raw", type=int, default=180000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = S.LangPretrainer(backbone=dict(LANG_BACKBONE), criteria=[
    dict(type="CosineSimilarity", reduction="mean", loss_weight=1.0), dict(type="L2Loss", reduction="mean", loss_weight=1.0),
    dict(type="AggregatedContrastiveLoss", temperature=0.2, reduction="mean", loss_weight=0.02, schedule="last_75")]).to(dev).train()
net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.05)

# one collated batch per rank: `chunks` GridSampled chunks concatenated with cumulative offsets (collate_fn)
gs = S.GridSample(grid_size=0.02, hash_type="fnv", mode="train", keys=("coord", "color", "opacity", "quat", "scale"),
                  return_grid_coord=True, device=dev)
parts, sizes = [], []
for c in range(args.chunks):
    d = synthetic.chunk(args.n_raw, seed=rank * 100 + c)
    np.random.seed(c)
    out = gs({k: torch.from_numpy(v) for k, v in d.items() if k in ("coord", "color", "opacity", "quat", "scale")})
    parts.append(out)
    sizes.append(out["coord"].shape[0])
n = sum(sizes)
g = torch.Generator().manual_seed(rank)
batch = dict(coord=torch.cat([p["coord"] for p in parts]), grid_coord=torch.cat([p["grid_coord"] for p in parts]),
             feat=torch.cat([torch.cat([p["color"], p["opacity"], p["quat"], p["scale"]], 1) for p in parts]).contiguous(),
             offset=torch.tensor(np.cumsum(sizes), device=dev),
             lang_feat=torch.nn.functional.normalize(torch.randn(n, 768, generator=g), dim=1).to(dev),
             valid_feat_mask=(torch.rand(n, generator=g) < 0.8).to(dev),
             segment=torch.randint(-1, 200, (n,), generator=g).to(dev), epoch_progress=0.9)


def step(timing=None):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    loss = net(dict(batch))["loss"]
    ev[1].record()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    ev[2].record()
    opt.step()
    ev[3].record()
    if timing is not None:
        timing.append(ev)
    return loss


for _ in range(args.warmup):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
timing = []
t0 = time.perf_counter()
for _ in range(args.steps):
    loss = step(timing)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = (time.perf_counter() - t0) / args.steps
tot = torch.tensor([float(n)], device=dev)
if world > 1:
    dist.all_reduce(tot)
if rank == 0:
    f = np.mean([e[0].elapsed_time(e[1]) for e in timing])
    b = np.mean([e[1].elapsed_time(e[2]) for e in timing])
    o = np.mean([e[2].elapsed_time(e[3]) for e in timing])
    print(f"train step: {world} GPU(s) x {args.chunks} chunks ({n} voxels on rank 0), loss {float(loss):.4f}: "
          f"{1e3 * dt:.1f} ms/step = {float(tot) / dt / 1e6:.2f} M Gaussians/s; rank 0 forward+loss {f:.1f} ms, "
          f"backward{' + all-reduce' if world > 1 else ''} {b:.1f} ms, AdamW {o:.1f} ms; "
          f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
if world > 1:
    dist.destroy_process_group()
