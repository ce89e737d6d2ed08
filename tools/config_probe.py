"""Timings of BASELINE.json's other configurations on one GPU (the bench line is configuration C2; these are informational):
C1 one 752x480 frame pair through the per-frame interfaces, C3 Aria RGB 1408x1408 / 4000 keypoints with mask (grid bucketing
off and on), C4 stereo pairs 640x480 / 1000 keypoints through the batched tracker (left -> right matching)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aria_slam_b200 import synth
from aria_slam_b200._lib import Context
from aria_slam_b200.frontend import OrbBatchTracker, OrbB200Extractor, B200Matcher, Frame

ctx = Context(0)
out = {}


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    ctx.sync(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


# C1: per-frame interfaces, host buffers
a = synth.frame(752, 480, 1); b = synth.shifted(a, seed=2)
ex, mt = OrbB200Extractor(1000, ctx), B200Matcher(ctx)
fa, fb = Frame(), Frame()
ex.extract(a, 752, 480, fa)
t_ex = timed(lambda: ex.extract(b, 752, 480, fb), 50)
mm = []
t_m = timed(lambda: mt.match(fa, fb, [], 0.75), 50)
mt.match(fa, fb, mm, 0.75)
out["C1"] = {"extract_ms": t_ex * 1e3, "match_ms": t_m * 1e3, "keypoints": fb.numKeypoints(), "matches": len(mm),
             "pairs_per_s": 1.0 / (t_ex + t_m)}
ex.close()
# C3: 1408x1408, 4000 keypoints, mask; batches of 16 frames, host buffers in and out
w = h = 1408
frames = np.stack([synth.frame(w, h, 3000 + i) for i in range(4)] * 4)
masks = np.stack([synth.mask_and_boxes(w, h, 3001 + i, n=6)[0] for i in range(4)] * 4)
for cell in (0, 64):
    tr = OrbBatchTracker(w, h, 4000, 16, ctx, grid_cell=cell)
    cap = tr.capacity
    kps = np.zeros((16, cap), dtype=[("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4")])
    desc = np.zeros((16, cap, 32), np.uint8); cnt = np.zeros(16, np.int32)

    def step():
        tr.extract_async(frames, masks)
        tr.sync(kps, desc, cnt)
    t = timed(step, 10)
    out["C3_grid_cell_%d" % cell] = {"frames_per_s": 16 / t, "ms_per_frame": t / 16 * 1e3, "keypoints_per_frame": float(cnt.mean())}
    tr.close()
# C4: stereo pairs 640x480, 1000 keypoints each, left -> right matching; 64 pairs per call (batch [L, R, L, R, ...])
w, h = 640, 480
pairs = []
for i in range(8):
    left = synth.frame(w, h, 4000 + i)
    pairs += [left, synth.shifted(left, dx=12, dy=0, sigma=1.0, seed=4100 + i)]
batch = np.stack(pairs * 8)                                    # 128 frames = 64 pairs
tr = OrbBatchTracker(w, h, 1000, len(batch), ctx)


def step4():
    tr.track_async(batch, 0.75)
    return tr.results(with_matches=True)
t = timed(step4, 10)
kps, desc, cnt, m, mc = step4()
out["C4"] = {"stereo_pairs_per_s": len(batch) / 2 / t, "ms_per_pair": t / (len(batch) / 2) * 1e3, "keypoints_per_frame": float(cnt.mean()),
             "left_right_matches_per_pair": float(mc[1::2].mean()), "realtime_30hz_cameras_per_gpu": len(batch) / 2 / t / 30.0}
tr.close()
print(json.dumps(out))
