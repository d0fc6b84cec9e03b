"""Correctness + timing of the CTA-pair Linear (+GELU) kernel (csrc/gemm2cta.cu) against cuBLASLt (+ the GELU kernel)
at MLP shapes of the lang config (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from scenesplat_b200 import ops, _lib
if os.environ.get("SS_LIB"):  # A/B against a variant build of the library (developer tool only)
    _lib.LIB_PATH = os.path.abspath(os.environ["SS_LIB"])

shapes = [(1000, 64, 256), (513, 768, 3072), (299277, 768, 3072), (299277, 3072, 768), (119000, 512, 2048), (299277, 768, 2304)]
if len(sys.argv) > 3:
    shapes = [tuple(int(v) for v in sys.argv[1:4])]
elif len(sys.argv) > 1 and sys.argv[1] == "all":   # every Linear shape of the lang config's Blocks (N per level of the bench chunk)
    shapes = []
    for n, c in ((299277, 32), (119000, 64), (36000, 128), (9100, 256), (36000, 256), (119000, 512), (299277, 768)):
        shapes += [(n, c, 3 * c), (n, c, c), (n, c, 4 * c), (n, 4 * c, c)]
reps = 5


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


for n, cin, cout in shapes:
    torch.manual_seed(0)
    x = torch.randn(n, cin, device="cuda").bfloat16()
    w = (torch.randn(cout, cin, device="cuda") / cin ** 0.5).bfloat16()
    b = torch.randn(cout, device="cuda")
    for act in (0, 1):
        got = ops.linear_act(x, w, b, act)
        torch.cuda.synchronize()
        ref = F.linear(x.float(), w.float(), b) if n <= 20000 else F.linear(x, w, b.bfloat16()).float()
        if act:
            ref = F.gelu(ref)
        err = (got.float() - ref).abs().max().item()
        rel = ((got.float() - ref).norm() / ref.norm()).item()
        t_own = timed(lambda: ops.linear_act(x, w, b, act))
        bb = b.bfloat16()
        t_lib = timed(lambda: ops.affine_act(F.linear(x, w, bb), act=1) if act else F.linear(x, w, bb))
        fl = 2.0 * n * cin * cout
        print(f"n={n} cin={cin} cout={cout} act={act}: max err {err:.4f} rel {rel:.2e} | own {t_own:.3f} ms "
              f"({fl / t_own / 1e9:.0f} TFLOP/s), cuBLASLt{' + GELU kernel' if act else ''} {t_lib:.3f} ms "
              f"({fl / t_lib / 1e9:.0f} TFLOP/s)", flush=True)


# ---- fc2 + bias + residual add in the epilogue (RES = 1) against cuBLASLt + the package's add kernel
print("residual epilogue (fp32 residual stream updated in place + bf16 copy):")
for n, cin, cout in [(299277, 3072, 768), (119000, 2048, 512), (36000, 1024, 256), (299277, 128, 32)]:
    torch.manual_seed(0)
    x = torch.randn(n, cin, device="cuda").bfloat16()
    w = (torch.randn(cout, cin, device="cuda") / cin ** 0.5).bfloat16()
    b = torch.randn(cout, device="cuda")
    res = torch.randn(n, cout, device="cuda")
    got, sh = ops.linear_residual(x, w, b, res.clone())
    ref = res + F.linear(x, w, b.bfloat16()).float()
    rel = ((got - ref).norm() / ref.norm()).item()
    r1 = res.clone()
    t_own = timed(lambda: ops.linear_residual(x, w, b, r1))
    bb = b.bfloat16()
    r2 = res.clone()
    t_lib = timed(lambda: ops.add_layernorm(r2, F.linear(x, w, bb), None, None, norm_dtype=torch.bfloat16, inplace=True))
    fl = 2.0 * n * cin * cout
    print(f"n={n} cin={cin} cout={cout}: rel {rel:.2e} | own fused {t_own:.3f} ms ({fl / t_own / 1e9:.0f} TFLOP/s), "
          f"cuBLASLt + add kernel {t_lib:.3f} ms", flush=True)
