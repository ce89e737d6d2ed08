"""Error behaviour of the C ABI (include/aslam_b200.h): every misuse is a status code + message, never a
crash or a silent fallback; results after an error-free call are unaffected by a previous failed one."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _params(**kw):
    from aria_slam_b200 import _lib
    p = _lib.OrbParams()
    _lib.lib().aslam_orb_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


@pytest.mark.parametrize("kw,w,h,batch", [
    (dict(nfeatures=-1), 640, 480, 1), (dict(nlevels=9), 640, 480, 1), (dict(nlevels=0), 640, 480, 1),
    (dict(scale_factor=1.0), 640, 480, 1), (dict(scale_factor=2.5), 640, 480, 1), (dict(edge_threshold=10), 640, 480, 1),
    (dict(wta_k=3), 640, 480, 1), (dict(patch_size=15), 640, 480, 1), (dict(fast_threshold=0), 640, 480, 1),
    (dict(grid_cell=7), 640, 480, 1), (dict(scale_factor=2.0, nlevels=3), 753, 480, 1),     # 753 -> 376: a factor above 2
    (dict(), 0, 480, 1), (dict(), 640, 5000, 1), (dict(), 640, 480, 0)])
def test_orb_create_rejects_bad_arguments(ctx, kw, w, h, batch):
    from aria_slam_b200 import _lib
    hnd = C.c_void_p()
    rc = _lib.lib().aslam_orb_create(ctx.h, C.byref(_params(**kw)), w, h, batch, C.byref(hnd))
    assert rc == _lib.E_INVALID and not hnd.value
    assert len(_lib.lib().aslam_last_error()) > 0


def test_call_order_violations_are_state_errors(ctx):
    from aria_slam_b200 import _lib, synth
    L = _lib.lib()
    hnd = C.c_void_p()
    assert L.aslam_orb_create(ctx.h, C.byref(_params(nfeatures=300)), 320, 240, 2, C.byref(hnd)) == _lib.OK
    cap = L.aslam_orb_capacity(hnd)
    kps = np.zeros((2, cap), _lib.KP_DTYPE); desc = np.zeros((2, cap, 32), np.uint8); cnt = np.zeros(2, np.int32)
    assert L.aslam_orb_sync(hnd, _lib.ptr(kps), _lib.ptr(desc), _lib.ptr(cnt)) == _lib.E_STATE        # nothing pending
    img = synth.frame(320, 240, 3)
    assert L.aslam_orb_extract_async(hnd, _lib.ptr(img), 320, 240, 320, None, None, 0) == _lib.OK
    assert L.aslam_orb_extract_async(hnd, _lib.ptr(img), 320, 240, 320, None, None, 0) == _lib.E_STATE  # already pending
    assert L.aslam_orb_set_max_features(hnd, 100) == _lib.E_STATE
    assert L.aslam_orb_sync(hnd, _lib.ptr(kps), _lib.ptr(desc), _lib.ptr(cnt)) == _lib.OK
    first = (int(cnt[0]), kps[0, :cnt[0]].copy(), desc[0, :cnt[0]].copy())
    assert L.aslam_orb_extract_async(hnd, _lib.ptr(img), 640, 240, 640, None, None, 0) == _lib.E_INVALID  # wrong frame size
    imgs = np.stack([img, img, img])
    assert L.aslam_orb_extract_batch_async(hnd, _lib.ptr(imgs), 3, 320 * 240, 320, None) == _lib.E_INVALID    # above max_batch
    # the handle still works and gives the same answer
    assert L.aslam_orb_extract_async(hnd, _lib.ptr(img), 320, 240, 320, None, None, 0) == _lib.OK
    assert L.aslam_orb_sync(hnd, _lib.ptr(kps), _lib.ptr(desc), _lib.ptr(cnt)) == _lib.OK
    assert int(cnt[0]) == first[0] and (kps[0, :cnt[0]] == first[1]).all() and (desc[0, :cnt[0]] == first[2]).all()
    L.aslam_orb_destroy(hnd)


def test_matcher_argument_errors_and_degenerate_sizes(ctx):
    from aria_slam_b200 import _lib
    L = _lib.lib()
    rs = np.random.RandomState(3)
    q = rs.randint(0, 256, (10, 32), dtype=np.uint8); t = rs.randint(0, 256, (1, 32), dtype=np.uint8)
    out = np.zeros(16, _lib.MATCH_DTYPE); n = C.c_int(-1)
    # one train row: no second neighbour, so no match (CudaMatcher.cpp:60) -- not an error
    assert L.aslam_match_knn2(ctx.h, _lib.ptr(q), 10, _lib.ptr(t), 1, C.c_float(0.75), _lib.ptr(out), C.byref(n)) == _lib.OK and n.value == 0
    assert L.aslam_match_knn2(ctx.h, _lib.ptr(q), 0, _lib.ptr(t), 1, C.c_float(0.75), _lib.ptr(out), C.byref(n)) == _lib.OK and n.value == 0
    t5 = rs.randint(0, 256, (5, 32), dtype=np.uint8)
    assert L.aslam_match_knn2(ctx.h, None, 10, _lib.ptr(t5), 5, C.c_float(0.75), _lib.ptr(out), C.byref(n)) == _lib.E_INVALID
    assert L.aslam_match_knn2(ctx.h, _lib.ptr(q), -1, _lib.ptr(t5), 5, C.c_float(0.75), _lib.ptr(out), C.byref(n)) == _lib.E_INVALID
    assert L.aslam_match_knn2(None, _lib.ptr(q), 10, _lib.ptr(t5), 5, C.c_float(0.75), _lib.ptr(out), C.byref(n)) == _lib.E_INVALID


def test_loopdb_grows_instead_of_refusing_and_a_failed_add_changes_nothing(ctx, oracle):
    """Like the reference's std::deque (LoopClosure.cpp:24-31) the store never refuses a keyframe: the array grows when the
    live rows outgrow the initial capacity, with eviction, slide and growth interleaved; queries stay exact throughout.
    A call that fails (bad argument) leaves the database as it was."""
    from aria_slam_b200 import _lib
    from aria_slam_b200.frontend import LoopDB
    rs = np.random.RandomState(4)
    db = LoopDB(ctx, 1000, 4)
    kfs = [rs.randint(0, 256, (n, 32), dtype=np.uint8) for n in (600, 1200, 300, 900, 2500, 40, 700)]
    db.add(kfs[0], 0)
    db.add(kfs[1], 1)                                  # one keyframe larger than the whole initial store
    assert db.num_keyframes == 2 and db.num_rows == 1800 and db.capacity_rows >= 1800
    for i in range(2, len(kfs)):
        db.add(kfs[i], i)
        live = kfs[max(0, i - 3):i + 1]                # FIFO cap 4
        assert db.num_keyframes == len(live) and db.num_rows == sum(len(k) for k in live)
        flat = np.concatenate(live)
        q = np.concatenate([live[0][:30], live[-1][:30]])
        start = np.concatenate([[0], np.cumsum([len(k) for k in live])]).astype(np.int32)
        assert (db.query_segmented(q, 0.7) == oracle.loop_segment_counts(q, flat, start, 7, 10)).all()
        rows, dist = db.query_global(q)
        ridx, rdist = oracle.knn2(q, flat)
        evicted = sum(len(k) for k in kfs[:max(0, i - 3)])          # global rows keep counting from creation
        assert (rows - evicted == ridx).all() and (dist == rdist).all()
    before = (db.num_keyframes, db.num_rows)
    with pytest.raises(_lib.AslamError) as e:
        _lib.check(_lib.lib().aslam_loopdb_add(db.h, None, 5, 99))
    assert e.value.code == _lib.E_INVALID and (db.num_keyframes, db.num_rows) == before
    db.close()


def test_a_refused_cluster_launch_falls_back_to_level_by_level():
    """Calls with few frames build the pyramid with one thread-block-cluster launch and record their chain into a CUDA graph.
    On a device that refuses the cluster launch (simulated: ASLAM_TEST_REFUSE_CLUSTER=1 in a child process) the recording
    fails, the chain is issued directly, the level-by-level launches take over for good -- and the results are the oracle's."""
    import os, subprocess, sys, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent("""
        import sys; sys.path.insert(0, %r)
        import numpy as np
        from aria_slam_b200 import synth
        from aria_slam_b200._lib import Context
        from aria_slam_b200.frontend import OrbB200Extractor, Frame
        from oracle import oracle as O
        ctx = Context(0)
        ex = OrbB200Extractor(600, ctx)
        img = synth.frame(640, 480, 77)
        key = lambda k: sorted(zip(k["octave"].tolist(), k["x"].tolist(), k["y"].tolist()))
        ko, do = O.orb(img, 600)
        for _ in range(3):
            f = Frame(); ex.extract(img, 640, 480, f)
            assert key(f.keypoints) == key(ko)
        print("fallback ok", f.numKeypoints())
    """ % root)
    env = dict(os.environ, ASLAM_TEST_REFUSE_CLUSTER="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "fallback ok" in r.stdout, r.stdout + r.stderr
