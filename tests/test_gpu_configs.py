"""The remaining BASELINE.json configurations as parity cases (C3: 1408x1408 with mask and
dynamic boxes through the batched path; C4: stereo 640x480 pair, left/right matching), plus
randomised sizes for the matcher and the resize chain (hypothesis)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

from conftest import kp_keys

pytestmark = pytest.mark.gpu


def test_c3_aria_rgb_batch_with_masks(ctx, oracle):
    from aria_slam_b200 import synth
    from aria_slam_b200.frontend import OrbBatchTracker
    w = h = 1408
    frames = np.stack([synth.frame(w, h, 3000 + i) for i in range(2)])
    masks = np.stack([synth.mask_and_boxes(w, h, 3000 + i)[0] for i in range(2)])
    tr = OrbBatchTracker(w, h, 4000, 2, ctx)
    tr.extract_async(frames, masks=masks)
    kps, desc, cnt, _, _ = tr.results()
    for i in range(2):
        ko, do = oracle.orb(frames[i], 4000, mask=masks[i])
        k, d = kps[i, :cnt[i]], desc[i, :cnt[i]]
        assert sorted(kp_keys(k)) == sorted(kp_keys(ko))
        o1, o2 = np.lexsort((k["x"], k["y"], k["octave"])), np.lexsort((ko["x"], ko["y"], ko["octave"]))
        assert (d[o1] == do[o2]).all(1).mean() >= 0.99
        assert (k["angle"][o1] == ko["angle"][o2]).all() and (k["response"][o1] == ko["response"][o2]).all()
    tr.close()


def test_c4_stereo_pair_left_right_matching(ctx, oracle):
    """Dual 640x480 cameras: both images in one batch, then left (query) against right (train)."""
    from aria_slam_b200 import synth
    from aria_slam_b200.frontend import OrbBatchTracker, B200Matcher, Frame
    left = synth.frame(640, 480, 4000)
    right = synth.shifted(left, dx=12, dy=0, sigma=1.0, seed=4001)
    tr = OrbBatchTracker(640, 480, 1000, 2, ctx)
    tr.extract_async(np.stack([left, right]))
    kps, desc, cnt, _, _ = tr.results()
    fl = Frame(keypoints=kps[0, :cnt[0]], descriptors=desc[0, :cnt[0]].reshape(-1))
    fr = Frame(keypoints=kps[1, :cnt[1]], descriptors=desc[1, :cnt[1]].reshape(-1))
    m = []
    B200Matcher(ctx).match(fl, fr, m, 0.75)
    dl, dr = oracle.orb(left, 1000)[1], oracle.orb(right, 1000)[1]
    ol = np.lexsort((kps[0, :cnt[0]]["x"], kps[0, :cnt[0]]["y"], kps[0, :cnt[0]]["octave"]))
    idx, dist = oracle.knn2(desc[0, :cnt[0]], desc[1, :cnt[1]])
    ref = oracle.ratio_matches(idx, dist, 3, 4)
    got = np.array([(x["query_idx"], x["train_idx"], int(x["distance"])) for x in m], np.int32).reshape(-1, 3)
    assert got.shape == ref.shape and (got == ref).all()
    # geometry: a surviving match on level 0 is displaced by the 12-px shift
    k0, k1 = kps[0, :cnt[0]], kps[1, :cnt[1]]
    lvl0 = [(a, b) for a, b, _ in got if k0["octave"][a] == 0 and k1["octave"][b] == 0]
    dx = np.array([k1["x"][b] - k0["x"][a] for a, b in lvl0])
    assert len(lvl0) > 50 and (np.abs(dx - 12) <= 1).mean() > 0.9
    assert len(dl) == cnt[0] and len(dr) == cnt[1]
    tr.close()


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(nq=st.integers(1, 700), nt=st.integers(0, 1500), seed=st.integers(0, 10_000), low=st.booleans())
def test_matcher_random_sizes(ctx, oracle, nq, nt, seed, low):
    from aria_slam_b200.frontend import B200Matcher
    rs = np.random.RandomState(seed)
    q = rs.randint(0, 256, (nq, 32), dtype=np.uint8); t = rs.randint(0, 256, (nt, 32), dtype=np.uint8)
    if low:
        q &= 3; t &= 3
    idx, dist = B200Matcher(ctx).knn2(q, t)
    ridx, rdist = oracle.knn2(q, t)
    assert (idx == ridx).all() and (dist == rdist).all()


@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(w=st.integers(80, 500), h=st.integers(80, 400), seed=st.integers(0, 1000))
def test_pyramid_and_orb_random_sizes(ctx, oracle, w, h, seed):
    """Ragged frame sizes: every built level bit-equal to the oracle's chain, and the same keypoints."""
    from aria_slam_b200.frontend import OrbBatchTracker
    rs = np.random.RandomState(seed)
    img = rs.randint(0, 256, (h, w), dtype=np.uint8)
    img[h // 4: h // 2, w // 4: w // 2] = 200
    tr = OrbBatchTracker(w, h, 300, 1, ctx)
    tr.extract_async(img[None])
    kps, desc, cnt, _, _ = tr.results()
    ref = oracle.pyramid(img)
    for l in range(8):
        if min(ref[l].shape) <= 62:
            break
        assert (tr.read_level(0, l) == ref[l]).all(), "level %d" % l
        assert (tr.read_level(0, l, blurred=True) == oracle.blur7(ref[l])).all(), "blurred level %d" % l
    ko, do = oracle.orb(img, 300)
    k, d = kps[0, :cnt[0]], desc[0, :cnt[0]]
    assert sorted(kp_keys(k)) == sorted(kp_keys(ko))
    o1, o2 = np.lexsort((k["x"], k["y"], k["octave"])), np.lexsort((ko["x"], ko["y"], ko["octave"]))
    assert (d[o1] == do[o2]).all()
    tr.close()


def _checker(w, h, s):
    y, x = np.mgrid[0:h, 0:w]
    return (((x // s) + (y // s)) % 2 * 200 + 20).astype(np.uint8)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["checker16", "checker40", "ramp", "lowcontrast", "saturated", "tiled94"])
def test_structured_frames_with_ties_match_oracle(ctx, oracle, kind):
    """Inputs with mass ties (equal FAST scores, equal Harris responses) and flat regions: the tie-inclusive
    retainBest cuts, the strict NMS and the empty levels must all follow cv::ORB."""
    from aria_slam_b200 import synth
    from aria_slam_b200.frontend import OrbB200Extractor, Frame
    w, h, n = 752, 480, 1000
    if kind.startswith("checker"):
        img = _checker(w, h, int(kind[7:]))
    elif kind == "ramp":
        y, x = np.mgrid[0:h, 0:w]
        img = ((x + 2 * y) % 256).astype(np.uint8)                 # sawtooth: long straight edges, no corners off the seams
    elif kind == "tiled94":
        # a 94x94 block repeated: identical corners tie at the level-0 cut, the frame returns MORE than nfeatures
        # keypoints (1013), which also exercises the second D2H pass for rows beyond nfeatures
        blk = synth.frame(w, h, 77)[:94, :94]
        img = np.tile(blk, (h // 94 + 1, w // 94 + 1))[:h, :w].copy()
    elif kind == "lowcontrast":
        img = (100 + synth.frame(w, h, 31) // 4).astype(np.uint8)   # contrast near the FAST threshold: levels below quota
    else:
        img = np.clip(synth.frame(w, h, 32).astype(np.int32) * 3, 0, 255).astype(np.uint8)   # large saturated areas
    ko, do = oracle.orb(img, n)
    ex = OrbB200Extractor(n, ctx)
    f = Frame()
    ex.extract(img, w, h, f)
    assert f.numKeypoints() == len(ko)
    if kind == "tiled94":
        assert len(ko) > n
    key = lambda k: sorted(zip(k["octave"].tolist(), k["x"].tolist(), k["y"].tolist()))
    assert key(f.keypoints) == key(ko)
    if len(ko):
        og = np.lexsort((f.keypoints["x"], f.keypoints["y"], f.keypoints["octave"]))
        oo = np.lexsort((ko["x"], ko["y"], ko["octave"]))
        assert (f.keypoints["response"][og] == ko["response"][oo]).all() and (f.keypoints["angle"][og] == ko["angle"][oo]).all()
        assert (f.descriptors.reshape(-1, 32)[og] == do[oo]).all(1).mean() >= 0.99
    ex.close()


@pytest.mark.parametrize("n,scale,nlevels,edge,thr", [(800, 1.5, 4, 31, 20), (1000, 2.0, 3, 31, 10), (600, 1.1, 8, 25, 40),
                                                      (1000, 1.2, 5, 40, 20), (500, 1.3, 6, 19, 7)])
def test_other_create_arguments_match_oracle(ctx, oracle, n, scale, nlevels, edge, thr):
    """cv::ORB::create arguments the reference never uses but the C ABI accepts: other scale factors (resize tables),
    level counts (quotas), edge and FAST thresholds.  Levels bit-exact, keypoints and descriptors as the oracle
    (which tests/test_oracle_cv2.py pins to cv2 at the same arguments)."""
    from aria_slam_b200 import synth
    from aria_slam_b200.frontend import OrbBatchTracker
    w, h = 752, 480
    img = synth.frame(w, h, 5)
    tr = OrbBatchTracker(w, h, n, 1, ctx, params=dict(scale_factor=scale, nlevels=nlevels, edge_threshold=edge,
                                                      fast_threshold=thr))
    tr.extract_async(img[None])
    kps, desc, cnt, _, _ = tr.results()
    Wl, Hl, _ = oracle.level_geometry(w, h, nlevels, scale)
    lev = img
    for l in range(1, nlevels):
        lev = oracle.resize_exact(lev, int(Wl[l]), int(Hl[l]))
        if Wl[l] > 2 * edge and Hl[l] > 2 * edge:
            assert (tr.read_level(0, l) == lev).all(), "level %d" % l
    ko, do = oracle.orb(img, n, nlevels=nlevels, scale=scale, edge=edge, fast_thr=thr)
    k, d = kps[0, :cnt[0]], desc[0, :cnt[0]]
    assert sorted(kp_keys(k)) == sorted(kp_keys(ko))
    o1, o2 = np.lexsort((k["x"], k["y"], k["octave"])), np.lexsort((ko["x"], ko["y"], ko["octave"]))
    for name in ("size", "angle", "response"):
        assert (k[name][o1] == ko[name][o2]).all(), name
    assert (d[o1] == do[o2]).all(1).mean() >= 0.99
    tr.close()
