"""Scene loading, reference style vs packed (developer tool): one synthetic scene folder with the
reference's asset files (`.npy` per attribute, lang_feat fp16 [N, 768]) is loaded to the GPU

  (a) the reference's way: np.load per asset + the get_data dtype conversions (pointcept/datasets/scannetgs.py:59-150),
      torch.from_numpy(...).cuda() per attribute (pageable memory: every copy synchronises);
  (b) packed: scene_io.load_scene (one readinto into one pinned buffer) + one non-blocking host-to-device copy.

Both read from the page cache (second pass)."""
import os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import scene_io as oio
from scenesplat_b200 import scene_io as sio

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
root = tempfile.mkdtemp(prefix="ss_loader_")
try:
    folder = os.path.join(root, "scene")
    os.makedirs(folder)
    rng = np.random.default_rng(0)
    assets = dict(coord=rng.random((n, 3), dtype=np.float32) * 8, color=rng.integers(0, 256, (n, 3)).astype(np.uint8),
                  opacity=rng.random(n, dtype=np.float32), quat=rng.standard_normal((n, 4), dtype=np.float32),
                  scale=rng.random((n, 3), dtype=np.float32), lang_feat=rng.standard_normal((n, 768), dtype=np.float32).astype(np.float16),
                  valid_feat_mask=(rng.random(n) < 0.8), segment20=rng.integers(-1, 20, n).astype(np.int64))
    for k, v in assets.items():
        np.save(os.path.join(folder, k + ".npy"), v)
    packed = os.path.join(root, "scene.sspk")
    size = sio.pack_scene(folder, packed)
    torch.zeros(1, device="cuda")

    def ref_way():
        d = oio.get_data(folder)
        return {k: torch.from_numpy(v).cuda() for k, v in d.items()}

    def packed_way():
        return sio.load_scene(packed).to_device("cuda")

    for name, fn in (("reference-style (np.load per asset + conversions + pageable copies)", ref_way),
                     ("packed (one pinned read + one copy)", packed_way)):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        t = sorted(ts)[len(ts) // 2]
        nbytes = sum(v.numel() * v.element_size() for v in out.values())
        print(f"{name}: {t * 1e3:.1f} ms per scene of {n} Gaussians ({nbytes / 1e6:.0f} MB on the device): "
              f"{nbytes / t / 1e9:.2f} GB/s, {n / t / 1e6:.2f} M Gaussians/s")
finally:
    shutil.rmtree(root, ignore_errors=True)
