"""Zero-shot inference on one synthetic scene, the whole flow of pointcept/engines/test.py:300-383 +
evaluator.py:785-834 on the GPU (BASELINE.json configs[2]: ~1.5 M Gaussians, 768-d lang head, K = 200 text
embeddings):  GridSample(0.02, mode="test") fragments -> LangPretrainer(chunk_size=600000) per fragment ->
pred[idx] += sigmoid(feat T^T) (fused head) -> max / argmax / threshold -> k = 25 neighbour voting ->
confusion-matrix update.  Random-init weights and a synthetic ground truth: the numbers are throughput, not accuracy.

  python tools/zero_shot_scene.py [n_gaussians] [room edge in m]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import scenesplat_b200 as S
from scenesplat_b200 import synthetic
from bench import LANG_BACKBONE

n_raw = int(sys.argv[1]) if len(sys.argv) > 1 else 1500000
L = float(sys.argv[2]) if len(sys.argv) > 2 else 14.0
K, ignore = 200, -1
dev = torch.device("cuda")
torch.manual_seed(0)
model = S.LangPretrainer(backbone=dict(LANG_BACKBONE), criteria=[]).eval().to(dev)
text = torch.nn.functional.normalize(torch.randn(K, 768, generator=torch.Generator().manual_seed(1)), dim=1).to(dev)
d = synthetic.chunk(n_raw, L=L, H=3.0, seed=0)
scene = {k: torch.from_numpy(v).to(dev) for k, v in d.items() if k in ("coord", "color", "opacity", "quat", "scale")}
gt = torch.from_numpy(((d["coord"][:, 0] * 1.3 + d["coord"][:, 1] * 0.7).astype(np.int64)) % K).to(dev)
gs = S.GridSample(grid_size=0.02, hash_type="fnv", mode="test", keys=("coord", "color", "opacity", "quat", "scale"),
                  return_grid_coord=True, device=dev)


def run():
    t = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    parts = gs(dict(scene))
    torch.cuda.synchronize(); t["gridsample_fragments"] = time.perf_counter() - t0
    pred = torch.zeros((n_raw, K), device=dev)
    t0 = time.perf_counter()
    n_fwd = 0
    with torch.no_grad():
        for p in parts:
            feat = torch.cat([p["color"], p["opacity"], p["quat"], p["scale"]], 1).contiguous()
            n = p["coord"].shape[0]
            inp = dict(coord=p["coord"], grid_coord=p["grid_coord"], feat=feat, offset=torch.tensor([n], device=dev))
            out = model(inp, chunk_size=600000)["point_feat"]["feat"]
            S.zero_shot_accumulate(pred, out, text, p["index"])
            n_fwd += n
    torch.cuda.synchronize(); t["fragments_forward_head"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    mx, arg = torch.max(pred, dim=1)
    arg[mx < 0.1] = ignore
    torch.cuda.synchronize(); t["argmax"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    voted = S.neighbor_voting(scene["coord"], arg, 25, ignore, K)
    torch.cuda.synchronize(); t["neighbor_voting_k25"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    conf = torch.zeros((K, K), dtype=torch.int64, device=dev)
    fn = torch.zeros(K, dtype=torch.int64, device=dev)
    S.confusion_update(gt, voted, K, ignore, conf, fn)
    torch.cuda.synchronize(); t["confusion"] = time.perf_counter() - t0
    return t, len(parts), n_fwd, int(conf.sum() + fn.sum())


run()  # warm-up (allocator, cuBLASLt heuristics)
t, nfrag, n_fwd, counted = run()
tot = sum(t.values())
print(f"scene: {n_raw} Gaussians, {nfrag} fragments, {n_fwd} voxels through the encoder, {counted} points counted")
for k, v in t.items():
    print(f"  {k:26s} {1e3 * v:9.1f} ms")
print(f"  total {1e3 * tot:.1f} ms = {n_raw / tot / 1e6:.2f} M scene Gaussians/s end to end "
      f"({n_fwd / t['fragments_forward_head'] / 1e6:.2f} M voxels/s through encoder + head)")
