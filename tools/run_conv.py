"""Stand-alone driver of the tcgen05 gather-GEMM conv at a lang-config shape (developer tool for timing / ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scenesplat_b200 import ops, synthetic, _lib
if os.environ.get("SS_LIB"):  # A/B against a variant build of the library (developer tool only)
    _lib.LIB_PATH = os.path.abspath(os.environ["SS_LIB"])
from oracle import gridsample as ogs, serialization as oser

n_raw = int(os.environ.get("CONV_NRAW", 360000))
c = int(os.environ.get("CONV_C", 768))
reps = int(os.environ.get("CONV_REPS", 5))
d = synthetic.chunk(n_raw, L=6.0, H=3.0, seed=0)
res = ogs.grid_sample_train(d["coord"], 0.02)
g = res["grid_coord"]
n = g.shape[0]
offset = np.array([n], dtype=np.int64)
batch = oser.offset2batch(offset)
code, order, inv, depth = oser.serialization(g, batch, 1, ("z",))
dev = lambda a: torch.as_tensor(a).cuda()
nbr, cnt = ops.kmap_build(dev(g), dev(batch), dev(code[0]), dev(order[0]), depth, 0, 3)
torch.manual_seed(0)
x = torch.randn(n, c, device="cuda").bfloat16()
w = (torch.randn(27, c, c, device="cuda") * 0.01).bfloat16()
b = torch.zeros(c, device="cuda")
ref = None
for tile, fn in ((256, "ss_subm_conv_gemm256"), (256, "ss_subm_conv_gemm_pair")):
    if fn == "ss_subm_conv_gemm_pair" and c < 256:
        continue
    pairs = ops.kmap_pairs(nbr, dev(order[0]), 3, cnt.cpu().numpy(), tile=tile)
    prod = torch.empty((pairs["p_pad"], c), dtype=torch.bfloat16, device="cuda")
    from scenesplat_b200 import _lib as L
    def run():
        L.call(fn, L.ptr(x), L.ptr(pairs["pair_in"]), L.ptr(w), L.ptr(pairs["tile_tap"]), pairs["p_pad"], 27, c, c,
               L.ptr(prod), L.stream())
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * pairs["pairs"] * c * c
    print(f"n={n} C={c} {fn}: pairs {pairs['pairs']} (padded {pairs['p_pad']}) gemm {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s useful")
    if ref is None:
        ref = prod.clone()
    else:
        print(f"   {fn} bit-identical to the single-CTA kernel:", bool(torch.equal(ref, prod)))

# ---- the whole xCPE stage of a Block: gather-GEMM + gather-sum + LN + residual + LN, two launches vs one
if c >= 256:
    pairs = ops.kmap_pairs(nbr, dev(order[0]), 3, cnt.cpu().numpy(), tile=256)
    resid = torch.randn(n, c, device="cuda")
    ln0 = (torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda"))
    ln1 = (torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda"))
    outs = {}
    for impl in ("split", "fused"):
        if os.environ.get("CONV_ONLY") and os.environ["CONV_ONLY"] != impl:
            continue
        for _ in range(2):
            y, h = ops.subm_conv_gemm_add_ln(x, pairs, w, b, resid, ln0, ln1, 1e-5, inplace=False, impl=impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            y, h = ops.subm_conv_gemm_add_ln(x, pairs, w, b, resid, ln0, ln1, 1e-5, inplace=False, impl=impl)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        outs[impl] = (y, h)
        print(f"n={n} C={c} conv + gather-sum + LN + residual + LN, {'two launches' if impl == 'split' else 'ONE launch (fused)'}: "
              f"{ms:.3f} ms  {2.0 * pairs['pairs'] * c * c / ms / 1e9:.0f} TFLOP/s useful")
    if len(outs) == 2:
        print("   fused bit-identical to split:", bool(torch.equal(outs["fused"][0], outs["split"][0]) and
                                                      torch.equal(outs["fused"][1], outs["split"][1])))
if os.environ.get("CONV_ONLY"):
    sys.exit(0)

# ---- backward kernels at the same shape: dgrad = the forward kernels on mirrored taps, wgrad = csrc/conv_wgrad.cu
from scenesplat_b200 import training
pairs = ops.kmap_pairs(nbr, dev(order[0]), 3, cnt.cpu().numpy(), tile=256)
dy = torch.randn(n, c, device="cuda").bfloat16()
po = training._pair_out(pairs)
for _ in range(2):
    dw = ops.subm_conv_wgrad(x, dy, pairs, po, 27)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    dw = ops.subm_conv_wgrad(x, dy, pairs, po, 27)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"n={n} C={c} wgrad: {ms:.3f} ms  {2.0 * pairs['pairs'] * c * c / ms / 1e9:.0f} TFLOP/s useful (incl. zeroing dW)")
