"""Device-resident frames/s of the 256-frame step cut into S frame streams (S from argv)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aria_slam_b200._lib import Context
from aria_slam_b200.frontend import OrbBatchTracker
from aria_slam_b200 import synth

B, W, H, N = 256, 752, 480, 2000
frames = synth.frame_stream(32, W, H, 1000)
d = torch.from_numpy(np.ascontiguousarray(np.concatenate([frames] * (B // 32)))).cuda()
main = torch.cuda.Stream()
out = {}
for S in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
    pipes = []
    for s in range(S):
        st = torch.cuda.Stream(); c = Context(0, st.cuda_stream)
        pipes.append((st, c, OrbBatchTracker(W, H, N, B // S, c)))
    def run(k):
        for _ in range(k):
            for s, (st, c, t) in enumerate(pipes):
                t.track_async(d[s * (B // S):(s + 1) * (B // S)], 0.75, on_device=True); t.device_results()
    run(3); torch.cuda.synchronize()
    K = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for st, c, t in pipes: st.wait_event(e0)
    run(K)
    for st, c, t in pipes:
        ev = torch.cuda.Event(); ev.record(st); main.wait_event(ev)
    e1.record(main); torch.cuda.synchronize()
    out[S] = round(B * K / (e0.elapsed_time(e1) * 1e-3))
    for st, c, t in pipes: t.close()
print(json.dumps(out))
