"""compute-sanitizer is closed on the GPU pool, so memory safety is checked by the library itself: the debug build
(libaslam_b200_dbg.so, -DASLAM_DEBUG_BOUNDS=1) checks every staged index of every kernel on the device and counts the
failures.  This test runs ragged frame sizes, masks, degenerate frames, ragged descriptor counts and the loop database
through that build (a subprocess, because the library is chosen at import time) and wants zero violations and results
identical to the release build's."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBG = os.path.join(ROOT, "aria_slam_b200", "libaslam_b200_dbg.so")

WORKLOAD = r'''
import ctypes as C, json, sys, zlib
sys.path.insert(0, %r)
import numpy as np
from aria_slam_b200 import synth, _lib
from aria_slam_b200._lib import Context, lib
from aria_slam_b200.frontend import OrbBatchTracker, OrbB200Extractor, B200Matcher, LoopDB, Frame
ctx = Context(0)
h = 0
for (w, hh, n, batch) in [(752, 480, 2000, 3), (333, 257, 700, 2), (100, 80, 300, 1), (1408, 1408, 4000, 1), (640, 480, 1000, 2), (67, 200, 100, 1)]:
    frames = np.stack([synth.frame(w, hh, 40 + i) if i %% 2 == 0 else synth.natural_frame(w, hh, 40 + i) for i in range(batch)])
    tr = OrbBatchTracker(w, hh, n, batch, ctx)
    tr.track_async(frames, 0.75)
    kps, desc, cnt, m, mc = tr.results(with_matches=True)
    for i in range(batch):
        h = zlib.crc32(kps[i, :cnt[i]].tobytes(), h); h = zlib.crc32(desc[i, :cnt[i]].tobytes(), h); h = zlib.crc32(m[i, :mc[i]].tobytes(), h)
    tr.close()
ex = OrbB200Extractor(1500, ctx)
img = synth.frame(752, 480, 77)
mask, boxes = synth.mask_and_boxes(752, 480, 78)
for kw in ({}, {"mask": mask}, {"boxes": boxes}):
    f = Frame(); ex.extract(img, 752, 480, f, **kw)
    h = zlib.crc32(f.keypoints.tobytes(), h); h = zlib.crc32(f.descriptors.tobytes(), h)
f0 = Frame(); ex.extract(np.full((480, 752), 9, np.uint8), 752, 480, f0)        # constant frame: no keypoints
ex.close()
mt = B200Matcher(ctx)
rs = np.random.RandomState(5)
for nq, nt in [(1, 2), (7, 129), (513, 1), (2000, 2001), (129, 3000)]:
    q = rs.randint(0, 256, (nq, 32), dtype=np.uint8); t = rs.randint(0, 256, (nt, 32), dtype=np.uint8)
    idx, dist = mt.knn2(q, t)
    h = zlib.crc32(idx.tobytes(), h); h = zlib.crc32(dist.tobytes(), h)
db = LoopDB(ctx, 3000, 7)
for k in range(12):
    db.add(rs.randint(0, 256, (rs.randint(1, 900), 32), dtype=np.uint8), k)
q = rs.randint(0, 256, (777, 32), dtype=np.uint8)
h = zlib.crc32(db.query_segmented(q, 0.7).tobytes(), h)
r, d = db.query_global(q)
h = zlib.crc32(r.tobytes(), h); h = zlib.crc32(d.tobytes(), h)
h = zlib.crc32(db.match_keyframe(3, q, 0.7).tobytes(), h)
v = C.c_uint64(0)
enabled = lib().aslam_debug_bounds(C.byref(v))
print(json.dumps({"enabled": enabled, "violations": v.value, "digest": h}))
'''


def _run(libpath):
    env = dict(os.environ)
    if libpath:
        env["ASLAM_B200_LIB"] = libpath
    else:
        env.pop("ASLAM_B200_LIB", None)
    r = subprocess.run([sys.executable, "-c", WORKLOAD % ROOT], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])


def test_debug_library_is_built_and_exports_the_abi():
    from aria_slam_b200 import _lib
    import ctypes as C
    assert os.path.exists(DBG), "run `make -C aria_slam_b200/csrc debug` (done by __graft_entry__.build())"
    l = C.CDLL(DBG)
    for name in _lib.SYMBOLS:
        getattr(l, name)


@pytest.mark.gpu
def test_no_staged_index_leaves_its_buffer_on_ragged_inputs():
    dbg = _run(DBG)
    assert dbg["enabled"] == 1 and dbg["violations"] == 0, dbg
    rel = _run(None)
    assert rel["enabled"] == 0 and rel["digest"] == dbg["digest"]
