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
