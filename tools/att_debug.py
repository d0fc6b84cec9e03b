"""Developer tool: tcgen05 attention vs the SIMT kernel, error map per (patch, query tile, head)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scenesplat_b200 import ops

def run(n, H, d, K, seed=0):
    torch.manual_seed(seed)
    C = H * d
    qkv = (torch.randn(n, 3 * C, device="cuda") * 1.5).bfloat16()
    order = torch.randperm(n, device="cuda")
    offset = torch.tensor([n], device="cuda")
    table = ops.patch_table(offset, K, n)
    a = ops.patch_attention(qkv, order, table, K, H, d ** -0.5, impl="tc").float()
    b = ops.patch_attention(qkv, order, table, K, H, d ** -0.5, impl="simt").float()
    torch.cuda.synchronize()
    err = (a - b).abs()[order]            # sorted positions
    err = err.view(n, H, d).amax(dim=2)   # [n, H]
    print(f"n={n} H={H} d={d} K={K}: max err {err.max().item():.4f}")
    bad = (err > 0.05)
    if bad.any():
        pos = torch.nonzero(bad)
        tiles = {}
        for p, h in pos.tolist():
            key = (p // K, (p % K) // 128, h)
            tiles[key] = tiles.get(key, 0) + 1
        for k in sorted(tiles)[:40]:
            print("   patch %d tile %d head %d: %d bad rows" % (*k, tiles[k]))

for cfg in [(1024, 1, 32, 1024), (1024, 1, 48, 1024), (1024, 2, 16, 1024), (256, 1, 16, 256), (512, 1, 16, 256), (256, 4, 16, 256),
            (2048, 3, 32, 1024), (300, 1, 32, 1024), (128, 1, 32, 1024), (129, 1, 32, 1024)]:
    run(*cfg)
