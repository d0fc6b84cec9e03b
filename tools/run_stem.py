"""Stand-alone driver of the stem conv (5^3, 11 -> 32, SIMT) at the lang-config size (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scenesplat_b200 import ops, synthetic
from oracle import gridsample as ogs, serialization as oser

n_raw = int(os.environ.get("CONV_NRAW", 360000))
reps = int(os.environ.get("CONV_REPS", 5))
d = synthetic.chunk(n_raw, L=6.0, H=3.0, seed=0)
res = ogs.grid_sample_train(d["coord"], 0.02)
g = res["grid_coord"]
n = g.shape[0]
offset = np.array([n], dtype=np.int64)
batch = oser.offset2batch(offset)
code, order, inv, depth = oser.serialization(g, batch, 1, ("z",))
dev = lambda a: torch.as_tensor(a).cuda()
nbr, cnt = ops.kmap_build(dev(g), dev(batch), dev(code[0]), dev(order[0]), depth, 0, 5)
torch.manual_seed(0)
x = torch.randn(n, 11, device="cuda")
w = torch.randn(125, 11, 32, device="cuda") * 0.05
sc, sh = torch.rand(32, device="cuda") + 0.5, torch.randn(32, device="cuda")
for _ in range(2):
    y = ops.subm_conv_simt(x, nbr, w, None, sc, sh, act=1, out_dtype=torch.float32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    y = ops.subm_conv_simt(x, nbr, w, None, sc, sh, act=1, out_dtype=torch.float32)
e1.record()
torch.cuda.synchronize()
print(f"n={n} stem conv 5^3 11->32: {e0.elapsed_time(e1) / reps:.3f} ms, active pairs {int(cnt.sum())} ({int(cnt.sum()) / n:.1f}/voxel)")
