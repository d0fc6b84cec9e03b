"""torch.profiler breakdown of one training step (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0], "--chunks", "1", "--n-raw", "180000", "--steps", "1", "--warmup", "2"]
import torch
from torch.profiler import profile, ProfilerActivity
import runpy
ns = runpy.run_path(os.path.join(os.path.dirname(__file__), "train_step.py"), run_name="not_main")
step = ns["step"]
step(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by=os.environ.get("PROF_SORT", "cuda_time_total"), row_limit=40, max_name_column_width=70))
