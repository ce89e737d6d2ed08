"""Where a single frame's latency goes: per-kernel event times of one-frame calls (752x480 / 1000, 1408x1408 / 4000)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aria_slam_b200._lib import Context
from aria_slam_b200.frontend import OrbBatchTracker
from aria_slam_b200 import synth

ctx = Context(0)
out = {}
for (w, h, n, B) in ((752, 480, 1000, 1), (752, 480, 2000, 1), (1408, 1408, 4000, 1), (752, 480, 2000, 4)):
    frames = synth.frame_stream(max(B, 2), w, h, 7)[:B]
    d = torch.from_numpy(np.ascontiguousarray(frames)).cuda()
    tr = OrbBatchTracker(w, h, n, B, ctx)
    for _ in range(5):
        tr.extract_async(d, on_device=True); tr.device_results()
    ctx.sync()
    K = 50
    t0 = time.perf_counter()
    for _ in range(K):
        tr.extract_async(d, on_device=True); tr.device_results(); ctx.sync()
    wall = (time.perf_counter() - t0) / K * 1e3
    ctx.prof_enable(True)
    for _ in range(K):
        tr.extract_async(d, on_device=True); tr.device_results()
    ctx.sync()
    st = {k: round(v[0] / K, 4) for k, v in ctx.prof_read(reset=True).items() if v[1]}
    ctx.prof_enable(False)
    out["%dx%d/%d x%d" % (w, h, n, B)] = {"wall_ms_device_input": round(wall, 4), "kernel_ms": st, "kernel_sum_ms": round(sum(st.values()), 4)}
    tr.close()
print(json.dumps(out, indent=1))
