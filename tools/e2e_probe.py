"""How the host-buffer path scales with the number of interleaved streams, and the raw PCIe rates."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aria_slam_b200._lib import Context
from aria_slam_b200.frontend import OrbBatchTracker
from aria_slam_b200 import synth

W, H, B = 752, 480, 256
frames = synth.frame_stream(32, W, H, 1000)
h_in = torch.from_numpy(np.ascontiguousarray(np.concatenate([frames] * (B // 32)))).pin_memory()
out = {}
# raw copies
d = torch.empty_like(h_in, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(h_in, non_blocking=True)), ("d2h", lambda: h_in.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize()
    out[name + "_GBps"] = h_in.numel() * 10 / (time.perf_counter() - t0) / 1e9
for ns in (1, 2, 4, 8):
    per = B // ns
    pipes = []
    for s in range(ns):
        st = torch.cuda.Stream()
        c = Context(0, st.cuda_stream)
        t = OrbBatchTracker(W, H, 2000, per, c)
        cap = t.capacity
        o = dict(kps=torch.empty((per, cap, 24), dtype=torch.uint8).pin_memory(), desc=torch.empty((per, cap, 32), dtype=torch.uint8).pin_memory(),
                 cnt=torch.empty(per, dtype=torch.int32).pin_memory(), m=torch.empty((per, cap, 12), dtype=torch.uint8).pin_memory(),
                 mc=torch.empty(per, dtype=torch.int32).pin_memory())
        pipes.append((c, t, o, h_in[s * per:(s + 1) * per]))
    def step():
        for c, t, o, src in pipes: t.track_async(src, 0.75, on_device=False)
        for c, t, o, src in pipes: t.sync(o["kps"], o["desc"], o["cnt"], o["m"], o["mc"])
    for _ in range(3): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); K = 10
    for _ in range(K): step()
    torch.cuda.synchronize()
    out["e2e_fps_streams_%d" % ns] = B * K / (time.perf_counter() - t0)
    for c, t, o, src in pipes: t.close()
print(json.dumps(out, indent=1))
