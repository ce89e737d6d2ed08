"""Oracle (CPU) and CUDA path (GPU) against the committed golden vectors made from cv2 by
tests/golden/make_golden.py."""
import os

import numpy as np
import pytest

from conftest import kp_keys

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _orb_check(k, d, gk, gd):
    assert sorted(kp_keys(k)) == sorted(kp_keys(gk))
    o1, o2 = np.lexsort((k["x"], k["y"], k["octave"])), np.lexsort((gk["x"], gk["y"], gk["octave"]))
    assert (k["response"][o1] == gk["response"][o2]).all() and (k["angle"][o1] == gk["angle"][o2]).all()
    assert (k["size"][o1] == gk["size"][o2]).all()
    return float((np.asarray(d).reshape(-1, 32)[o1] == gd[o2]).all(1).mean())


def test_oracle_reproduces_golden_orb(oracle):
    g = np.load(os.path.join(G, "orb_320x240.npz"))
    k, d = oracle.orb(g["image"], 500)
    assert _orb_check(k, d, g["keypoints"], g["descriptors"]) == 1.0
    k, d = oracle.orb(g["image"], 500, mask=g["mask"])
    assert _orb_check(k, d, g["keypoints_masked"], g["descriptors_masked"]) == 1.0
    assert (oracle.resize_exact(g["image"], 267, 200) == g["level1"]).all()
    c = oracle.fast_nms(oracle.fast_score_map(g["image"], 20))
    assert sorted(zip(c["y"].tolist(), c["x"].tolist(), c["score"].tolist())) == [tuple(r) for r in g["fast_yxs"].tolist()]


def test_oracle_reproduces_golden_knn(oracle):
    g = np.load(os.path.join(G, "knn_ties.npz"))
    idx, dist = oracle.knn2(g["query"], g["train"])
    assert (idx == g["idx"]).all() and (dist == g["dist"]).all()
    assert (oracle.ratio_matches(idx, dist, 3, 4)[:, 0] == g["pass_075"]).all()
    assert (oracle.ratio_matches(idx, dist, 7, 10)[:, 0] == g["pass_070"]).all()


@pytest.mark.gpu
def test_gpu_reproduces_golden_orb(ctx):
    from aria_slam_b200.frontend import OrbB200Extractor, Frame
    g = np.load(os.path.join(G, "orb_320x240.npz"))
    ex = OrbB200Extractor(500, ctx)
    f = Frame(); ex.extract(g["image"], 320, 240, f)
    assert _orb_check(f.keypoints, f.descriptors, g["keypoints"], g["descriptors"]) >= 0.99
    f = Frame(); ex.extract(g["image"], 320, 240, f, mask=g["mask"])
    assert _orb_check(f.keypoints, f.descriptors, g["keypoints_masked"], g["descriptors_masked"]) >= 0.99
    ex.close()


@pytest.mark.gpu
def test_gpu_reproduces_golden_knn(ctx):
    from aria_slam_b200.frontend import B200Matcher, Frame
    g = np.load(os.path.join(G, "knn_ties.npz"))
    mt = B200Matcher(ctx)
    idx, dist = mt.knn2(g["query"], g["train"])
    assert (idx == g["idx"]).all() and (dist == g["dist"]).all()
    m = []
    mt.match(Frame(descriptors=g["query"].reshape(-1)), Frame(descriptors=g["train"].reshape(-1)), m, 0.75)
    assert [int(x["query_idx"]) for x in m] == g["pass_075"].tolist()


def _config_frames():
    import zlib
    from aria_slam_b200 import synth
    g = np.load(os.path.join(G, "orb_configs.npz"))
    for name in ("c1", "c2", "c4"):
        w, h, n, seed, crc = (int(v) for v in g[name + "_spec"])
        img = synth.frame(w, h, seed)
        assert zlib.crc32(img.tobytes()) == crc, "the synthetic frame generator drifted"
        yield name, img, w, h, n, g[name + "_keypoints"], g[name + "_descriptors"]


def test_oracle_reproduces_golden_configs(oracle):
    """cv2's answers for the BASELINE configurations C1, C2 and C4 at their real sizes."""
    for name, img, w, h, n, gk, gd in _config_frames():
        k, d = oracle.orb(img, n)
        assert _orb_check(k, d, gk, gd) == 1.0, name


@pytest.mark.gpu
def test_gpu_reproduces_golden_configs(ctx):
    from aria_slam_b200.frontend import OrbB200Extractor, Frame
    for name, img, w, h, n, gk, gd in _config_frames():
        ex = OrbB200Extractor(n, ctx)
        f = Frame(); ex.extract(img, w, h, f)
        assert _orb_check(f.keypoints, f.descriptors, gk, gd) >= 0.99, name
        ex.close()


def _c3():
    import zlib
    from aria_slam_b200 import synth
    g = np.load(os.path.join(G, "orb_c3.npz"))
    w, h, n, seed, mseed, crc, mcrc = (int(v) for v in g["spec"])
    img = synth.frame(w, h, seed)
    mask, boxes = synth.mask_and_boxes(w, h, mseed, n=6)
    assert zlib.crc32(img.tobytes()) == crc and zlib.crc32(mask.tobytes()) == mcrc, "the synthetic generators drifted"
    return img, mask, w, h, n, g["keypoints"], g["descriptors"]


def test_oracle_reproduces_golden_c3(oracle):
    """cv2's answer for configuration C3: 1408x1408, 4000 keypoints, dynamic-object mask."""
    img, mask, w, h, n, gk, gd = _c3()
    k, d = oracle.orb(img, n, mask=mask)
    assert _orb_check(k, d, gk, gd) == 1.0


def _c5():
    import zlib
    from aria_slam_b200 import synth
    g = np.load(os.path.join(G, "loop_c5.npz"))
    rows, dseed, nq, qseed, dcrc, qcrc = (int(v) for v in g["spec"])
    db = synth.descriptor_db(rows, dseed)
    q, src = synth.loop_queries(db, nq, qseed)
    assert zlib.crc32(db[:100_000].tobytes()) == dcrc and zlib.crc32(q.tobytes()) == qcrc, "the synthetic generators drifted"
    return g, db, q, src


def test_oracle_reproduces_golden_c5_on_a_sample(oracle):
    """cv2's neighbours over the whole 1 M-row database (all 4000 keys are in the fixture; the scalar oracle checks 24 queries)
    and its per-keyframe ratio-0.7 counts on the sampled keyframes."""
    import zlib
    from aria_slam_b200.sharded import pack_top2
    g, db, q, src = _c5()
    assert zlib.crc32(g["keys"].tobytes()) == int(g["keys_crc32"])
    pick = np.arange(0, len(q), len(q) // 24)[:24]
    idx, dist = oracle.knn2(q[pick], db)
    assert (pack_top2(idx, dist) == g["keys"][pick]).all()
    for kf, want in zip(g["keyframe_sample"], g["good_sample"]):
        got = oracle.loop_segment_counts(q, db[kf * 2000:(kf + 1) * 2000], np.array([0, 2000], np.int32), 7, 10)
        assert int(got[0]) == int(want), kf


@pytest.mark.gpu
def test_gpu_reproduces_golden_c3(ctx):
    from aria_slam_b200.frontend import OrbB200Extractor, Frame
    img, mask, w, h, n, gk, gd = _c3()
    ex = OrbB200Extractor(n, ctx)
    f = Frame(); ex.extract(img, w, h, f, mask=mask)
    assert _orb_check(f.keypoints, f.descriptors, gk, gd) >= 0.99
    ex.close()


@pytest.mark.gpu
def test_gpu_reproduces_golden_c5_every_key(ctx):
    """All 4000 neighbour keys (both neighbours and both distances of all 2000 queries) of the 1 M-row query equal cv2's,
    and so do the per-keyframe counts on the sampled keyframes."""
    import zlib
    from aria_slam_b200.frontend import LoopDB
    from aria_slam_b200._lib import check, lib, ptr
    g, db, q, src = _c5()
    whole = LoopDB(ctx, len(db), 500)
    for k in range(500):
        whole.add(db[k * 2000:(k + 1) * 2000], k)
    keys = np.zeros((len(q), 2), np.uint64)
    check(lib().aslam_loopdb_query_global(whole.h, ptr(q), len(q), ptr(keys)))
    assert (keys == g["keys"]).all() and zlib.crc32(keys.tobytes()) == int(g["keys_crc32"])
    good = whole.query_segmented(q, 0.7)
    assert (good[g["keyframe_sample"]] == g["good_sample"]).all()
    whole.close()
