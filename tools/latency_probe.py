"""Single-frame latency through the interface calls (host image in, host results out): C1-like use."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aria_slam_b200._lib import Context
from aria_slam_b200.frontend import OrbB200Extractor, B200Matcher, Frame
from aria_slam_b200 import synth

ctx = Context(0)
out = {}
for (w, h, n) in ((752, 480, 1000), (752, 480, 2000), (640, 480, 1000), (1408, 1408, 4000)):
    a = synth.frame(w, h, 1); b = synth.shifted(a, seed=2)
    ex = OrbB200Extractor(n, ctx); m = B200Matcher(ctx)
    fa, fb = Frame(), Frame()
    for _ in range(5):
        ex.extract(a, w, h, fa); ex.extract(b, w, h, fb); mm = []; m.match(fb, fa, mm, 0.75)
    K = 50
    t0 = time.perf_counter()
    for _ in range(K):
        ex.extract(b, w, h, fb)
    t1 = time.perf_counter()
    for _ in range(K):
        mm = []; m.match(fb, fa, mm, 0.75)
    t2 = time.perf_counter()
    out["%dx%d/%d" % (w, h, n)] = {"extract_ms": (t1 - t0) / K * 1e3, "match_ms": (t2 - t1) / K * 1e3, "kp": fb.numKeypoints(), "matches": len(mm)}
    ex.close()
print(json.dumps(out, indent=1))
