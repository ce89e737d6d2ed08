"""Small fixed workload for ncu: 2 steps (1 warm-up + 1 profiled) of the batched extract+match
path on 256 frames, then one loop query.  Run plain first, then under ncu (B200_PROFILING.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aria_slam_b200._lib import Context, kernel_launches
from aria_slam_b200.frontend import OrbBatchTracker
from aria_slam_b200 import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = Context(0)
frames = synth.frame_stream(32, 752, 480, 1000)
d = torch.from_numpy(np.ascontiguousarray(np.concatenate([frames] * (B // 32)))).cuda()
tr = OrbBatchTracker(752, 480, 2000, B, ctx)
for it in range(2):
    tr.track_async(d, 0.75, on_device=True); tr.device_results()
    ctx.sync()
    print("step", it, "launches so far", kernel_launches())
