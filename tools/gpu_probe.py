"""First-contact probe on the B200 box: integer-pipe micro-benchmarks + a timing of each stage."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aria_slam_b200._lib import Context
from aria_slam_b200.frontend import OrbBatchTracker
from aria_slam_b200 import synth

ctx = Context(0)
out = {"sm_count": ctx.sm_count}
for which, name in enumerate(["popc", "lop3", "iadd3", "vimnmx3_u16x2", "hmnmx2", "vimnmx3+hmnmx2", "imad", "prmt", "ffma", "ffma2_lanes", "fadd2_lanes", "fmul2_lanes"]):
    out["gops_" + name] = ctx.microbench(which)
if os.environ.get("ASLAM_PROBE_FRAMES", "") == "natural":       # 1/f noise: the FAST quick-reject discards most pixel pairs
    frames = np.stack([synth.natural_frame(752, 480, 1000 + i) if i % 2 == 0 else synth.shifted(synth.natural_frame(752, 480, 999 + i), seed=i)
                       for i in range(32)])
else:
    frames = synth.frame_stream(32, 752, 480, 1000)
B = 256
batch = np.ascontiguousarray(np.concatenate([frames] * (B // 32)))
import torch
d = torch.from_numpy(batch).cuda()
tr = OrbBatchTracker(752, 480, 2000, B, ctx, grid_cell=int(os.environ.get("ASLAM_PROBE_GRID_CELL", "0")))
ctx.prof_enable(True)
for it in range(3):
    tr.track_async(d, 0.75, on_device=True); tr.device_results()
ctx.sync(); ctx.prof_read(reset=True)
t0 = time.perf_counter()
K = 5
for it in range(K):
    tr.track_async(d, 0.75, on_device=True); tr.device_results()
ctx.sync()
dt = (time.perf_counter() - t0) / K
out["ms_per_batch256"] = dt * 1e3
out["frames_per_s_device"] = B / dt
out["stages_ms_per_batch"] = {k: (v[0] / K, v[1] / K) for k, v in ctx.prof_read().items()}
ctx.prof_enable(False)
for it in range(2):
    tr.track_async(d, 0.75, on_device=True); tr.device_results()
ctx.sync()
t0 = time.perf_counter()
for it in range(K):
    tr.track_async(d, 0.75, on_device=True); tr.device_results()
ctx.sync()
out["ms_per_batch256_noprof"] = (time.perf_counter() - t0) / K * 1e3
print(json.dumps(out, indent=1))
