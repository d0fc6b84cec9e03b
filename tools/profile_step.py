"""Where does a bench step spend host time?  (developer tool; not part of the product path)"""
import cProfile, pstats, sys, os, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
model = bench.build_model().to(dev)
dev_in, host_in, n = bench.voxelize_on_gpu(bench.make_chunk(0), dev)
def step():
    with torch.no_grad():
        return model(dict(dev_in))["point_feat"]["feat"]
for _ in range(3):
    step()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"enqueue {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms")
pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35); print(s.getvalue()[:6000])
