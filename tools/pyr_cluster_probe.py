"""Wall time of an extract call (device input, device results) against the frames per call; run once with
ASLAM_PYR_CLUSTER_MAX=0 (level-by-level launches) and once with ASLAM_PYR_CLUSTER_MAX=4096 (cluster kernel always)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aria_slam_b200._lib import Context
from aria_slam_b200.frontend import OrbBatchTracker
from aria_slam_b200 import synth

ctx = Context(0)
out = {"ASLAM_PYR_CLUSTER_MAX": os.environ.get("ASLAM_PYR_CLUSTER_MAX"), "ASLAM_NO_GRAPH": os.environ.get("ASLAM_NO_GRAPH")}
base = synth.frame_stream(32, 752, 480, 1000)
for B in (1, 2, 4, 8, 16, 32, 64, 256):
    d = torch.from_numpy(np.ascontiguousarray(np.concatenate([base] * ((B + 31) // 32))[:B])).cuda()
    tr = OrbBatchTracker(752, 480, 2000, B, ctx)
    for _ in range(5):
        tr.extract_async(d, on_device=True); tr.device_results()
    ctx.sync()
    K = 40 if B <= 64 else 10
    t0 = time.perf_counter()
    for _ in range(K):
        tr.extract_async(d, on_device=True); tr.device_results(); ctx.sync()
    wall = (time.perf_counter() - t0) / K * 1e3
    ctx.prof_enable(True)
    for _ in range(K):
        tr.extract_async(d, on_device=True); tr.device_results()
    ctx.sync()
    st = ctx.prof_read(reset=True)
    ctx.prof_enable(False)
    out[str(B)] = {"extract_wall_ms": round(wall, 4), "resize_ms_events": round(st["resize"][0] / K, 4), "select_ms_events": round(st["select_harris"][0] / K, 4)}
    tr.close()
print(json.dumps(out))
