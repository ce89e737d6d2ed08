"""Per-level grid bucketing of the Harris cull (aslam_orb_params::grid_cell) -- an extension BASELINE.json's
north_star names; neither cv::ORB nor the reference buckets, so PARITY IS UNPINNED here:
the CPU tests pin the C restatement to an independent numpy statement of the definition and to its
properties, the GPU tests compare the CUDA path with the C restatement.  With grid_cell = 0 everything is
cv::ORB (all other test files)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

from conftest import kp_keys


def _numpy_bucket_keep(xs, ys, resp, quota, cell):
    """The definition, stated independently: rank = number of same-cell corners with a greater response;
    sort by (rank, -response); keep the first `quota` and all that tie with the quota-th on both."""
    n = len(xs)
    if quota <= 0:
        return np.zeros(n, bool)
    if n <= quota:
        return np.ones(n, bool)
    cellid = (np.asarray(ys) // cell).astype(np.int64) * 100000 + (np.asarray(xs) // cell)
    resp = np.asarray(resp, np.float32)
    rank = np.zeros(n, np.int64)
    for c in np.unique(cellid):
        idx = np.nonzero(cellid == c)[0]
        r = resp[idx]
        rank[idx] = (r[None, :] > r[:, None]).sum(1)
    order = np.lexsort((-resp.astype(np.float64), rank))
    cut_rank, cut_resp = rank[order[quota - 1]], resp[order[quota - 1]]
    return (rank < cut_rank) | ((rank == cut_rank) & (resp >= cut_resp))


@pytest.mark.parametrize("n,quota,cell,levels", [(900, 434, 32, None), (900, 434, 64, 7), (300, 60, 16, 3),
                                                 (50, 60, 32, None), (500, 1, 32, 2), (500, 499, 8, 5)])
def test_oracle_bucketing_equals_the_numpy_definition(oracle, n, quota, cell, levels):
    rng = np.random.RandomState(n + quota + cell)
    xs = rng.randint(31, 721, n); ys = rng.randint(31, 449, n)
    resp = rng.rand(n).astype(np.float32)
    if levels:                                   # few distinct responses: mass ties inside cells and at the cut
        resp = (np.floor(resp * levels) / levels).astype(np.float32)
    got = oracle.grid_bucket_keep(xs, ys, resp, quota, cell)
    ref = _numpy_bucket_keep(xs, ys, resp, quota, cell)
    assert (got == ref).all()
    assert got.sum() >= min(n, quota)            # the quota is always filled


def test_oracle_bucketing_properties(oracle):
    rng = np.random.RandomState(7)
    n, quota, cell = 800, 200, 64
    xs = rng.randint(31, 721, n); ys = rng.randint(31, 449, n)
    resp = rng.rand(n).astype(np.float32)          # distinct with probability ~1
    keep = oracle.grid_bucket_keep(xs, ys, resp, quota, cell)
    assert keep.sum() == quota
    cells = (ys // cell) * 1000 + xs // cell
    # every occupied cell gives its best corner before any cell gives its second (quota >= occupied cells here)
    assert quota >= len(np.unique(cells))
    for c in np.unique(cells):
        idx = np.nonzero(cells == c)[0]
        assert keep[idx[np.argmax(resp[idx])]]
        kept, dropped = resp[idx][keep[idx]], resp[idx][~keep[idx]]
        assert len(dropped) == 0 or kept.min() > dropped.max()      # a cell keeps a prefix of its own ranking
    # per-cell counts differ by at most one between cells that still had corners to give
    cnt = {c: int(keep[cells == c].sum()) for c in np.unique(cells)}
    full = max(cnt.values())
    assert all(v >= full - 1 or v == int((cells == c).sum()) for c, v in cnt.items())
    # one cell covering the level = cv::ORB's plain retainBest
    one = oracle.grid_bucket_keep(xs, ys, resp, quota, 4096)
    thr = np.sort(resp)[::-1][quota - 1]
    assert (one == (resp >= thr)).all()


def test_oracle_orb_with_bucketing_spreads_keypoints(oracle):
    from aria_slam_b200 import synth
    img = synth.frame(752, 480, 1)
    k0, _ = oracle.orb(img, 1000)
    k1, d1 = oracle.orb(img, 1000, grid_cell=32)
    assert len(k1) == len(k0) == 1000 and d1.shape == (1000, 32)
    occupied = lambda k: len({(int(x) // 64, int(y) // 64) for x, y in zip(k["x"][k["octave"] == 0], k["y"][k["octave"] == 0])})
    assert occupied(k1) > occupied(k0)
    # candidates are cv::ORB's (same first cull): the bucketed set is drawn from the same Harris-scored pool
    assert (np.bincount(k1["octave"], minlength=8) == np.bincount(k0["octave"], minlength=8)).all()


# ----------------------------------------------------------------------------- CUDA path vs the restatement
def _same(f_kps, f_desc, ko, do):
    assert sorted(kp_keys(f_kps)) == sorted(kp_keys(ko))
    og = np.lexsort((f_kps["x"], f_kps["y"], f_kps["octave"]))
    oo = np.lexsort((ko["x"], ko["y"], ko["octave"]))
    assert (f_kps["response"][og] == ko["response"][oo]).all() and (f_kps["angle"][og] == ko["angle"][oo]).all()
    assert (f_desc.reshape(-1, 32)[og] == do[oo]).all(1).mean() >= 0.99
    assert (np.diff(f_kps["octave"]) >= 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,seed,cell", [(752, 480, 1000, 1, 32), (752, 480, 2000, 2, 64), (752, 480, 1000, 3, 16),
                                             (640, 480, 1000, 4000, 8), (320, 240, 500, 11, 48), (752, 480, 1000, 1, 4096)])
def test_gpu_bucketed_orb_matches_the_restatement(ctx, oracle, w, h, n, seed, cell):
    from aria_slam_b200 import synth
    from aria_slam_b200.frontend import OrbB200Extractor, Frame
    img = synth.frame(w, h, seed)
    ex = OrbB200Extractor(n, ctx, grid_cell=cell)
    f = Frame()
    ex.extract(img, w, h, f)
    ko, do = oracle.orb(img, n, grid_cell=cell)
    _same(f.keypoints, f.descriptors, ko, do)
    if cell == 4096:                                  # one cell: cv::ORB itself
        kc, _ = oracle.orb(img, n)
        assert sorted(kp_keys(kc)) == sorted(kp_keys(f.keypoints))
    # switching it off on the same extractor gives cv::ORB again
    ex.setGridBucketing(0)
    ex.extract(img, w, h, f)
    kc, dc = oracle.orb(img, n)
    _same(f.keypoints, f.descriptors, kc, dc)
    ex.close()


@pytest.mark.gpu
@settings(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(w=st.integers(120, 520), h=st.integers(120, 420), n=st.integers(200, 1500), cell=st.integers(8, 120),
       seed=st.integers(0, 1000), natural=st.booleans())
def test_gpu_bucketed_random_sizes_and_cells(ctx, oracle, w, h, n, cell, seed, natural):
    """Ragged frame sizes, quotas and cell sizes, dense and sparse frames: levels over and under quota, cells cut by
    the level border, one or many cells per level."""
    from aria_slam_b200 import synth
    from aria_slam_b200.frontend import OrbBatchTracker
    img = synth.natural_frame(w, h, seed, 1.2) if natural else synth.frame(w, h, seed)
    tr = OrbBatchTracker(w, h, n, 1, ctx, grid_cell=cell)
    tr.extract_async(img[None])
    kps, desc, cnt, _, _ = tr.results()
    tr.close()
    ko, do = oracle.orb(img, n, grid_cell=cell)
    assert cnt[0] == len(ko)
    if len(ko):
        _same(kps[0, :cnt[0]], desc[0, :cnt[0]], ko, do)


@pytest.mark.gpu
def test_gpu_c3_bucketed_with_masks_batched(ctx, oracle):
    """BASELINE configuration C3 as worded: 1408x1408, 4000 keypoints, grid bucketing + mask, batched path."""
    from aria_slam_b200 import synth
    from aria_slam_b200.frontend import OrbBatchTracker
    w = h = 1408
    frames = np.stack([synth.frame(w, h, 3000 + i) for i in range(2)])
    masks = np.stack([synth.mask_and_boxes(w, h, 3000 + i)[0] for i in range(2)])
    tr = OrbBatchTracker(w, h, 4000, 2, ctx, grid_cell=64)
    tr.extract_async(frames, masks=masks)
    kps, desc, cnt, _, _ = tr.results()
    for i in range(2):
        ko, do = oracle.orb(frames[i], 4000, mask=masks[i], grid_cell=64)
        _same(kps[i, :cnt[i]], desc[i, :cnt[i]], ko, do)
    tr.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["checker16", "tiled94", "lowcontrast"])
def test_gpu_bucketed_structured_frames_with_ties(ctx, oracle, kind):
    """Mass ties in the in-cell ranks and at the cut (identical corners in every cell), levels below quota."""
    from aria_slam_b200 import synth
    from aria_slam_b200.frontend import OrbB200Extractor, Frame
    w, h, n = 752, 480, 1000
    if kind == "checker16":
        y, x = np.mgrid[0:h, 0:w]
        img = (((x // 16) + (y // 16)) % 2 * 200 + 20).astype(np.uint8)
    elif kind == "tiled94":
        blk = synth.frame(w, h, 77)[:94, :94]
        img = np.tile(blk, (h // 94 + 1, w // 94 + 1))[:h, :w].copy()
    else:
        img = (100 + synth.frame(w, h, 31) // 4).astype(np.uint8)
    for cell in (32, 47):
        ko, do = oracle.orb(img, n, grid_cell=cell)
        ex = OrbB200Extractor(n, ctx, grid_cell=cell)
        f = Frame()
        ex.extract(img, w, h, f)
        assert f.numKeypoints() == len(ko)          # tiled94 / cell 47 returns 1021 > nfeatures: ties at the cut are kept
        _same(f.keypoints, f.descriptors, ko, do)
        ex.close()


@pytest.mark.gpu
def test_gpu_grid_cell_argument_errors(ctx):
    from aria_slam_b200 import _lib
    import ctypes as C
    for bad in (-1, 1, 7, 5000):
        p = _lib.OrbParams()
        _lib.lib().aslam_orb_default_params(C.byref(p))
        p.grid_cell = bad
        h = C.c_void_p()
        assert _lib.lib().aslam_orb_create(ctx.h, C.byref(p), 320, 240, 1, C.byref(h)) == 1      # ASLAM_E_INVALID
        assert b"grid_cell" in _lib.lib().aslam_last_error()
