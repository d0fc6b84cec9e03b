"""Timing of the tcgen05 patch-attention backward (csrc/attention_bwd.cu) next to the library recomputation path
(torch SDPA in sorted space, scenesplat_b200/training.py::_attention_backward) at lang-config shapes (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scenesplat_b200 import ops
from scenesplat_b200 import training as T

shapes = [(163814, 16, 48), (65000, 16, 32), (163814, 2, 16), (20000, 16, 16)]
if len(sys.argv) > 3:
    shapes = [(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))]
K = 1024
reps = int(os.environ.get("ATT_REPS", 5))
lib = os.environ.get("ATT_LIB", "1") == "1"


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n, H, d in shapes:
    torch.manual_seed(0)
    C = H * d
    qkv = torch.randn(n, 3 * C, device="cuda").bfloat16()
    dout = torch.randn(n, C, device="cuda").bfloat16()
    order = torch.randperm(n, device="cuda")
    inverse = torch.empty_like(order)
    inverse[order] = torch.arange(n, device="cuda")
    offset = torch.tensor([n], device="cuda")
    table = ops.patch_table(offset, K, n)
    scale = d ** -0.5
    out, lse2 = ops.patch_attention_lse(qkv, order, table, K, H, scale)
    fwd = timed(lambda: ops.patch_attention_lse(qkv, order, table, K, H, scale))
    own = timed(lambda: ops.patch_attention_backward(qkv, out, dout, lse2, order, table, K, H, scale))
    msg = f"n={n} H={H} d={d}: forward {fwd:.3f} ms, own backward {own:.3f} ms ({own / fwd:.2f} x forward)"
    if lib:
        plan = [(0, n // K, n - (n // K) * K, n)]
        t = timed(lambda: T._attention_backward(qkv, order, inverse, plan, K, H, scale, dout))
        got = ops.patch_attention_backward(qkv, out, dout, lse2, order, table, K, H, scale).float()
        ref = T._attention_backward(qkv, order, inverse, plan, K, H, scale, dout).float()
        rel = float((got - ref).norm() / ref.norm())
        msg += f", SDPA recomputation path {t:.3f} ms; rel L2 own vs library {rel:.4f}"
    print(msg, flush=True)
