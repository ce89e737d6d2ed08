#!/usr/bin/env python
"""Generates the golden vectors of tests/golden/ from the live cv2 wheel -- the OpenCV code the
reference's CPU path calls (src/legacy/Frame.cpp:47, src/legacy/LoopClosure.cpp:21,87).

The reference repository holds no tests or vectors for this path, so these
fixtures are what pins the oracle (and, through it, the CUDA path) when cv2 is not consulted live.
Run from the repository root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
from aria_slam_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # ORB on a 320x240 frame, 500 features (small enough to commit: 77 KB of pixels)
    img = synth.frame(320, 240, 123)
    kps, desc = O.cv2_orb(img, 500)
    order = np.lexsort((kps["x"], kps["y"], kps["octave"]))
    mask, _ = synth.mask_and_boxes(320, 240, 124, n=5)
    kpm, descm = O.cv2_orb(img, 500, mask=mask)
    om = np.lexsort((kpm["x"], kpm["y"], kpm["octave"]))
    lvl = [cv2.resize(img, (267, 200), interpolation=cv2.INTER_LINEAR_EXACT)]
    fast = cv2.FastFeatureDetector_create(20, True).detect(img)
    fast = np.array(sorted((int(k.pt[1]), int(k.pt[0]), int(k.response)) for k in fast), np.int32)
    np.savez_compressed(os.path.join(HERE, "orb_320x240.npz"), image=img, keypoints=kps[order], descriptors=desc[order],
                        mask=mask, keypoints_masked=kpm[om], descriptors_masked=descm[om], level1=lvl[0], fast_yxs=fast,
                        cv2_version=np.array(cv2.__version__))
    # kNN: low-entropy descriptors with duplicates (ties), plus the ratio survivors at 0.75 and 0.7
    rs = np.random.RandomState(321)
    t = (rs.randint(0, 256, (400, 32)) % 4).astype(np.uint8)
    t[40:44] = t[9]
    q = np.concatenate([t[9:10], t[200:260], (rs.randint(0, 256, (60, 32)) % 4).astype(np.uint8),
                        rs.randint(0, 256, (40, 32)).astype(np.uint8)])
    q[70:90, 3] ^= 0x11
    idx, dist = O.cv2_knn2(q, t)
    r75 = np.array([i for i in range(len(q)) if np.float32(dist[i, 0]) < np.float32(0.75) * np.float32(dist[i, 1])], np.int32)
    r70 = np.array([i for i in range(len(q)) if float(dist[i, 0]) < 0.7 * float(dist[i, 1])], np.int32)
    np.savez_compressed(os.path.join(HERE, "knn_ties.npz"), query=q, train=t, idx=idx, dist=dist, pass_075=r75, pass_070=r70)
    # The BASELINE.json configurations at their real sizes: the frames come from the seeded generator (their
    # CRC32 is stored to catch generator drift), only cv2's keypoints and descriptors are committed.
    import zlib
    out = {}
    for name, (w, h, n, seed) in {"c1": (752, 480, 1000, 1), "c2": (752, 480, 2000, 1000), "c4": (640, 480, 1000, 4000)}.items():
        f = synth.frame(w, h, seed)
        k, d = O.cv2_orb(f, n)
        o = np.lexsort((k["x"], k["y"], k["octave"]))
        out[name + "_spec"] = np.array([w, h, n, seed, zlib.crc32(f.tobytes())], np.int64)
        out[name + "_keypoints"] = k[o]
        out[name + "_descriptors"] = d[o]
    np.savez_compressed(os.path.join(HERE, "orb_configs.npz"), cv2_version=np.array(cv2.__version__), **out)
    print("wrote orb_320x240.npz (%d kp, %d masked), knn_ties.npz and orb_configs.npz" % (len(kps), len(kpm)))


if __name__ == "__main__":
    main()
