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
    main_c3_c5()


def main_c3_c5():
    """Configuration C3 (Aria RGB 1408x1408 as grey, 4000 keypoints, dynamic-object mask) and C5 (2000 queries against the
    1 M-row keyframe database: all 4000 neighbour keys and the per-keyframe ratio-0.7 counts), again from cv2 itself."""
    import zlib
    w = h = 1408
    f = synth.frame(w, h, 3000)
    mask, _ = synth.mask_and_boxes(w, h, 3001, n=6)
    k, d = O.cv2_orb(f, 4000, mask=mask)
    o = np.lexsort((k["x"], k["y"], k["octave"]))
    np.savez_compressed(os.path.join(HERE, "orb_c3.npz"), cv2_version=np.array(cv2.__version__),
                        spec=np.array([w, h, 4000, 3000, 3001, zlib.crc32(f.tobytes()), zlib.crc32(mask.tobytes())], np.int64),
                        keypoints=k[o], descriptors=d[o])
    from aria_slam_b200.sharded import pack_top2
    db = synth.descriptor_db(1_000_000, 5)
    q, src = synth.loop_queries(db, 2000, 6)
    cv2.setNumThreads(os.cpu_count() or 1)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    idx, dist = O.cv2_knn2(q, db)                # in blocks of 250 000 rows: cv2 refuses 2^18 train rows or more
    keys = pack_top2(idx, dist)
    good = np.zeros(500, np.int32)               # findCandidates' counts (LoopClosure.cpp:86-95) for a sample of keyframes
    sample = sorted(set([0, 1, 7, 250, 499] + [int(r) // 2000 for r in src[:11]]))
    for kf in sample:
        good[kf] = sum(1 for m in bf.knnMatch(q, db[kf * 2000:(kf + 1) * 2000], k=2)
                       if len(m) >= 2 and m[0].distance < 0.7 * m[1].distance)
    np.savez_compressed(os.path.join(HERE, "loop_c5.npz"), cv2_version=np.array(cv2.__version__),
                        spec=np.array([1_000_000, 5, 2000, 6, zlib.crc32(db[:100_000].tobytes()), zlib.crc32(q.tobytes())], np.int64),
                        keys=keys, keys_crc32=np.array(zlib.crc32(keys.tobytes()), np.int64),
                        keyframe_sample=np.array(sample, np.int32), good_sample=good[sample])
    print("wrote orb_c3.npz (%d kp) and loop_c5.npz (crc %08x, %d keyframes sampled)" % (len(k), zlib.crc32(keys.tobytes()), len(sample)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "c3c5":
        main_c3_c5()
    else:
        main()
