"""Stand-alone driver of the tcgen05 patch-attention kernel at the dec0 / dec1 / enc0 shapes of the lang
config (developer tool for timing and ncu captures).

  python tools/run_attention.py                   # own kernel at ATT_N / ATT_H / ATT_D (env), default dec0
  python tools/run_attention.py --baseline flash  # + flash-attn 2.8.3 varlen with the reference's two row gathers
                                                  #   (ptv3:188,208-216) on the same tensors, three shapes
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from scenesplat_b200 import ops

if "--baseline" in sys.argv and sys.argv[sys.argv.index("--baseline") + 1] == "flash":
    import workloads as W
    for r in W.attention_library_bar(torch.device("cuda")):
        print(r)
    sys.exit(0)

n = int(os.environ.get("ATT_N", 299277))
H = int(os.environ.get("ATT_H", 16))
d = int(os.environ.get("ATT_D", 48))
K = 1024
reps = int(os.environ.get("ATT_REPS", 5))
torch.manual_seed(0)
C = H * d
qkv = torch.randn(n, 3 * C, device="cuda").bfloat16()
order = torch.randperm(n, device="cuda")
offset = torch.tensor([n], device="cuda")
table = ops.patch_table(offset, K, n)
for _ in range(2):
    out = ops.patch_attention(qkv, order, table, K, H, d ** -0.5, impl="tc")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = ops.patch_attention(qkv, order, table, K, H, d ** -0.5, impl="tc")
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flops = 4.0 * K * C * n
exps = n * K * H
print(f"n={n} H={H} d={d}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  {exps / ms / 1e9:.2f} Texp/s "
      f"(MUFU peak 148*16*1.965e9 = 4.65 T/s)")
