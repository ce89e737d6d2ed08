"""ctypes front-end of the CPU oracle (oracle/orb_oracle.c) plus the live cv2 reference calls.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (aria_slam_b200/) never imports it.

Parity status: the restatement is pinned against cv2 4.13.0 -- live in tests/test_oracle_cv2.py
and through the golden vectors in tests/golden/ (made by tests/golden/make_golden.py).  The
reference repository itself has no tests or golden vectors for this path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liborb_oracle.so")


class OrcParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int),
                ("edge_threshold", C.c_int), ("fast_threshold", C.c_int), ("grid_cell", C.c_int)]


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4")])
CORNER_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("score", "<i4")])

# 2 = the fused/unfused column split measured on this image's cv2 build (orb_oracle.c, A8);
# tests/test_oracle_cv2.py::test_blur_matches_cv2_sepfilter checks it.
BLUR_MODE = 2


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("orb_oracle.c", "orb_oracle.h", "brief_pattern_31.inc")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborb_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_harris.restype = C.c_float
        _lib.orc_fast_atan2.restype = C.c_float
        _lib.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.orc_ic_angle.restype = C.c_float
        _lib.orc_brief.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def level_geometry(W, H, nlevels=8, scale=1.2):
    Wl = np.zeros(nlevels, np.int32); Hl = np.zeros(nlevels, np.int32); sf = np.zeros(nlevels, np.float32)
    lib().orc_level_geometry(W, H, nlevels, C.c_float(scale), _p(Wl), _p(Hl), _p(sf))
    return Wl, Hl, sf


def level_quotas(nfeatures, nlevels=8, scale=1.2):
    N = np.zeros(nlevels, np.int32)
    lib().orc_level_quotas(nfeatures, C.c_float(scale), nlevels, _p(N))
    return N


def resize_coeffs(src, dst):
    ofs = np.zeros(dst, np.int32); c1 = np.zeros(dst, np.int32)
    lib().orc_resize_coeffs(src, dst, _p(ofs), _p(c1))
    return ofs, c1


def resize_exact(img, dw, dh):
    img = _u8(img)
    out = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_linear_exact(_p(img), img.shape[1], img.shape[0], img.shape[1], _p(out), dw, dh, dw)
    return out


def pyramid(img, nlevels=8, scale=1.2):
    img = _u8(img)
    Wl, Hl, _ = level_geometry(img.shape[1], img.shape[0], nlevels, scale)
    out = [img]
    for l in range(1, nlevels):
        out.append(resize_exact(out[-1], int(Wl[l]), int(Hl[l])))
    return out


def mask_pyramid(mask, nlevels=8, scale=1.2):
    """detectAndCompute(): level 0 as given; level l = threshold(resize(level l-1), 254, TOZERO)."""
    mask = _u8(mask)
    Wl, Hl, _ = level_geometry(mask.shape[1], mask.shape[0], nlevels, scale)
    out = [mask]
    for l in range(1, nlevels):
        m = resize_exact(out[-1], int(Wl[l]), int(Hl[l]))
        out.append(np.where(m > 254, m, 0).astype(np.uint8))
    return out


def fast_score_map(img, thr=20):
    img = _u8(img)
    s = np.zeros(img.shape, np.uint8)
    lib().orc_fast_score_map(_p(img), img.shape[1], img.shape[0], img.shape[1], thr, _p(s))
    return s


def fast_nms(score):
    score = _u8(score)
    cap = score.size // 4 + 16
    out = np.zeros(cap, CORNER_DTYPE)
    n = lib().orc_fast_nms(_p(score), score.shape[1], score.shape[0], _p(out), cap)
    return out[:n]


def grid_bucket_keep(xs, ys, resp, quota, cell):
    """Grid-bucketed second cull (extension, orb_oracle.c): boolean keep mask over the corners."""
    cs = np.zeros(len(xs), CORNER_DTYPE); cs["x"] = xs; cs["y"] = ys
    resp = np.ascontiguousarray(resp, np.float32)
    keep = np.zeros(len(xs) + 1, np.uint8)
    lib().orc_grid_bucket_keep(_p(cs), _p(resp), len(xs), int(quota), int(cell), _p(keep))
    return keep[:len(xs)].astype(bool)


def harris(img, x, y):
    img = _u8(img)
    return np.float32(lib().orc_harris(_p(img), img.shape[1], int(x), int(y)))


def fast_atan2(y, x):
    return np.float32(lib().orc_fast_atan2(float(np.float32(y)), float(np.float32(x))))


def ic_angle(img, x, y):
    img = _u8(img)
    return np.float32(lib().orc_ic_angle(_p(img), img.shape[1], int(x), int(y)))


def blur7(img, mode=None):
    img = _u8(img)
    out = np.zeros_like(img)
    lib().orc_blur7(_p(img), img.shape[1], img.shape[0], img.shape[1], _p(out), img.shape[1],
                    BLUR_MODE if mode is None else mode)
    return out


def blur7_split(img, row_fused_end, col_fused_end):
    """blur7 with explicit split points: fused multiply-adds for columns below row_fused_end (row pass) / col_fused_end
    (column pass), separately rounded products beyond."""
    img = _u8(img)
    out = np.zeros_like(img)
    lib().orc_blur7_split(_p(img), img.shape[1], img.shape[0], img.shape[1], _p(out), img.shape[1], int(row_fused_end), int(col_fused_end))
    return out


def brief(blurred, xi, yi, angle):
    blurred = _u8(blurred)
    d = np.zeros(32, np.uint8)
    lib().orc_brief(_p(blurred), blurred.shape[1], int(xi), int(yi), float(np.float32(angle)), _p(d))
    return d


def orb(img, nfeatures=1000, mask=None, nlevels=8, scale=1.2, edge=31, fast_thr=20, blur_mode=None, cap=None,
        grid_cell=0):
    """Whole ORB; returns (keypoints[KP_DTYPE], descriptors[n,32]) ordered by (level, y, x).
    grid_cell > 0 switches the second cull to the grid-bucketed extension (orb_oracle.c; not cv::ORB)."""
    img = _u8(img)
    if mask is not None:
        mask = _u8(mask)
    cap = cap or (4 * nfeatures + 4096)
    kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
    prm = OrcParams(nfeatures, scale, nlevels, edge, fast_thr, int(grid_cell))
    n = lib().orc_orb_detect_and_compute(_p(img), img.shape[1], img.shape[0], img.shape[1],
                                         _p(mask) if mask is not None else None, C.byref(prm),
                                         BLUR_MODE if blur_mode is None else blur_mode,
                                         _p(kps), _p(desc), cap)
    if n < 0:
        raise RuntimeError("oracle keypoint capacity exceeded")
    return kps[:n].copy(), desc[:n].copy()


def knn2(q, t):
    """Returns (idx[nq,2], dist[nq,2]); missing neighbours are -1."""
    q = _u8(q).reshape(-1, 32); t = _u8(t).reshape(-1, 32)
    idx = np.zeros((len(q), 2), np.int32); dist = np.zeros((len(q), 2), np.int32)
    lib().orc_knn2_hamming(_p(q), len(q), _p(t), len(t), _p(idx), _p(dist))
    return idx, dist


def ratio_matches(idx, dist, num, den):
    """CudaMatcher.cpp:59-67 on integers: rows (query, train, distance) that pass d0 < num/den*d1."""
    ok = (idx[:, 1] >= 0) & (den * dist[:, 0] < num * dist[:, 1])
    q = np.nonzero(ok)[0]
    return np.stack([q, idx[q, 0], dist[q, 0]], 1).astype(np.int32)


def loop_segment_counts(q, db, seg_start, num=7, den=10):
    q = _u8(q).reshape(-1, 32); db = _u8(db).reshape(-1, 32)
    seg_start = np.ascontiguousarray(seg_start, np.int32)
    good = np.zeros(len(seg_start) - 1, np.int32)
    lib().orc_loop_segment_counts(_p(q), len(q), _p(db), _p(seg_start), len(good), num, den, _p(good))
    return good


def loop_candidates(good, nq_keypoints, kf_ids, query_id, min_gap, top=5):
    """LoopClosure.cpp:72-114: score = good / max(1, #query keypoints); keep > 0.1 and outside the
    gap; sort by score descending (ties: lower keyframe index, our documented choice); top 5."""
    out = []
    for i, g in enumerate(good):
        if query_id - int(kf_ids[i]) < min_gap:
            continue
        s = float(g) / max(1, nq_keypoints)
        if s > 0.1:
            out.append((i, s))
    out.sort(key=lambda t: (-t[1], t[0]))
    return out[:top]


def merge_top2(parts):
    """Multi-GPU merge rule: parts = list of (idx[nq,2] global rows, dist[nq,2]); the two
    smallest (dist, idx) per query over all parts."""
    idx = np.concatenate([p[0] for p in parts], 1).astype(np.int64)
    dist = np.concatenate([p[1] for p in parts], 1).astype(np.int64)
    key = np.where(idx >= 0, dist * (1 << 32) + idx, np.iinfo(np.int64).max)
    order = np.argsort(key, 1, kind="stable")[:, :2]
    rows = np.arange(len(idx))[:, None]
    oi = idx[rows, order]; od = dist[rows, order]
    bad = key[rows, order] == np.iinfo(np.int64).max
    oi[bad] = -1; od[bad] = -1
    return oi.astype(np.int32), od.astype(np.int32)


# ---------------------------------------------------------------- live cv2 (the true reference)
def cv2_orb(img, nfeatures=1000, mask=None):
    """cv::ORB::create(n,1.2f,8,31,0,2,HARRIS_SCORE,31,20).detectAndCompute -- Frame.cpp:47 with
    the parameter set of OrbCudaExtractor.cpp:35-45."""
    import cv2
    o = cv2.ORB_create(nfeatures, 1.2, 8, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 20)
    k, d = o.detectAndCompute(np.ascontiguousarray(img), mask)
    kps = np.zeros(len(k), KP_DTYPE)
    for i, p in enumerate(k):
        kps[i] = (p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave)
    return kps, (d if d is not None else np.zeros((0, 32), np.uint8))


def cv2_knn2(q, t):
    """cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) -- LoopClosure.cpp:21,87.  cv2 refuses train sets of 2^18 rows or more
    (IMGIDX_ONE, matchers.cpp:860); larger sets are matched in blocks of 250 000 rows and the blocks' answers merged in
    (distance, row) order, which is the order cv2 itself uses inside a block."""
    import cv2
    q = _u8(q).reshape(-1, 32); t = _u8(t).reshape(-1, 32)
    idx = -np.ones((len(q), 2), np.int32); dist = -np.ones((len(q), 2), np.int32)
    if len(q) == 0 or len(t) == 0:
        return idx, dist
    if len(t) >= (1 << 18):
        parts = []
        for lo in range(0, len(t), 250_000):
            i, d = cv2_knn2(q, t[lo:lo + 250_000])
            parts.append((np.where(i >= 0, i.astype(np.int64) + lo, -1), d.astype(np.int64)))
        mi, md = merge_top2(parts)
        return mi.astype(np.int32), md.astype(np.int32)
    for i, m in enumerate(cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)):
        for j, e in enumerate(m):
            idx[i, j] = e.trainIdx; dist[i, j] = int(e.distance)
    return idx, dist
