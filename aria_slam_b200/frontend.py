"""Python mirror of the reference's frontend ports over the C ABI.

Same names, argument meaning and error behaviour as
  aria::interfaces::IFeatureExtractor  (include/interfaces/IFeatureExtractor.hpp:9-40)
  aria::interfaces::IMatcher           (include/interfaces/IMatcher.hpp:10-38)
  aria::interfaces::ILoopDetector      (include/interfaces/ILoopDetector.hpp:10-29)
so the parity tests read like tests of the reference's adapters.  The C++ adapters in
aria_slam_b200/adapters/ are the compiled-host-language form of the same thing; both are thin
layers over include/aslam_b200.h and neither computes anything on the CPU.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, MATCH_DTYPE, BOX_DTYPE, Context, check, lib, ptr


# ----------------------------------------------------------------------------- domain types
@dataclass
class Frame:
    """aria::core::Frame (include/core/Types.hpp:18-31): keypoints + flat N*32 descriptors."""
    id: int = 0
    timestamp: float = 0.0
    width: int = 0
    height: int = 0
    keypoints: np.ndarray = field(default_factory=lambda: np.zeros(0, KP_DTYPE))
    descriptors: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint8))

    def numKeypoints(self):
        return len(self.keypoints)

    def descriptorSize(self):
        return 32


@dataclass
class KeyFrame:
    """aria::core::KeyFrame (include/core/Types.hpp:34-45), descriptor-relevant part."""
    id: int = 0
    timestamp: float = 0.0
    frame: Frame = field(default_factory=Frame)


@dataclass
class LoopCandidate:
    """aria::core::LoopCandidate (include/core/Types.hpp:115-121) without the pose."""
    query_id: int
    match_id: int
    score: float
    matches: np.ndarray


def as_boxes(boxes):
    """Detections as aslam_box records: a BOX_DTYPE array, or rows of (x1, y1, x2, y2, confidence, class_id)."""
    boxes = np.asarray(boxes)
    if boxes.dtype == BOX_DTYPE:
        return np.ascontiguousarray(boxes).reshape(-1)
    rows = boxes.reshape(-1, 6)
    out = np.zeros(len(rows), BOX_DTYPE)
    for i, name in enumerate(("x1", "y1", "x2", "y2", "confidence")):
        out[name] = rows[:, i]
    out["class_id"] = rows[:, 5].astype(np.int32)
    return out


# ----------------------------------------------------------------------------- extractor
class OrbB200Extractor:
    """IFeatureExtractor over aslam_orb_* -- drop-in for OrbCudaExtractor
    (src/adapters/gpu/OrbCudaExtractor.cpp).  Results follow the reference's CPU path."""

    def __init__(self, max_features=1000, ctx: Optional[Context] = None, blur_mode=2, grid_cell=0):
        self.ctx = ctx or Context(0)
        self._owns_ctx = ctx is None
        self._n = int(max_features)
        self._blur_mode = blur_mode
        self._grid_cell = int(grid_cell)      # aslam_orb_params::grid_cell; 0 = cv::ORB's plain retainBest
        self._h = None
        self._size = None
        self._pending = None

    def _params(self):
        p = _lib.OrbParams()
        lib().aslam_orb_default_params(C.byref(p))
        p.nfeatures = self._n
        p.blur_mode = self._blur_mode
        p.grid_cell = self._grid_cell
        return p

    def _handle(self, width, height):
        if self._size != (width, height):
            if self._h:
                lib().aslam_orb_destroy(self._h)
                self._h = None
            h = C.c_void_p()
            p = self._params()
            check(lib().aslam_orb_create(self.ctx.h, C.byref(p), width, height, 1, C.byref(h)))
            self._h, self._size = h, (width, height)
        return self._h

    # IFeatureExtractor::extractAsync (IFeatureExtractor.hpp:28-33): returns at once, results after sync()
    def extractAsync(self, image_data, width, height, frame: Frame, mask=None, boxes=None):
        img = np.ascontiguousarray(image_data, np.uint8).reshape(height, width)
        h = self._handle(width, height)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8).reshape(height, width)
        b = None if boxes is None or len(boxes) == 0 else as_boxes(boxes)
        check(lib().aslam_orb_extract_async(h, ptr(img), width, height, width, ptr(m), ptr(b), 0 if b is None else len(b)))
        self._pending = (frame, width, height, img, m, b)    # keep the pixel buffers alive until sync()

    # IFeatureExtractor::sync (IFeatureExtractor.hpp:36)
    def sync(self):
        if self._pending is None:
            return
        frame, width, height = self._pending[:3]
        cap = lib().aslam_orb_capacity(self._h)
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8); cnt = np.zeros(1, np.int32)
        rc = lib().aslam_orb_sync(self._h, ptr(kps), ptr(desc), ptr(cnt))
        self._pending = None
        check(rc, allow_overflow=True)
        n = int(cnt[0])
        # OrbCudaExtractor.cpp:109-127 fills width/height/keypoints/descriptors only
        frame.width, frame.height = width, height
        frame.keypoints = kps[:n].copy()
        frame.descriptors = desc[:n].reshape(-1).copy()

    # IFeatureExtractor::extract (IFeatureExtractor.hpp:18-23)
    def extract(self, image_data, width, height, frame: Frame, mask=None, boxes=None):
        self.extractAsync(image_data, width, height, frame, mask, boxes)
        self.sync()

    def setMaxFeatures(self, n):
        self._n = int(n)
        if self._h:
            check(lib().aslam_orb_set_max_features(self._h, self._n))

    def getMaxFeatures(self):
        return self._n

    def setGridBucketing(self, cell_pixels):
        """Per-level grid bucketing of the Harris cull (extension, 0 = off); applies from the next extract."""
        if self._pending is not None:
            raise RuntimeError("setGridBucketing with an extract pending")
        if int(cell_pixels) != self._grid_cell:
            self._grid_cell = int(cell_pixels)
            if self._h:
                lib().aslam_orb_destroy(self._h)
                self._h, self._size = None, None

    def close(self):
        if self._h:
            lib().aslam_orb_destroy(self._h)
            self._h = None
        if self._owns_ctx:
            self.ctx.close()


# ----------------------------------------------------------------------------- matcher
class B200Matcher:
    """IMatcher over aslam_match_* -- drop-in for CudaMatcher (src/adapters/gpu/CudaMatcher.cpp)."""

    def __init__(self, ctx: Optional[Context] = None):
        self.ctx = ctx or Context(0)

    # IMatcher::match (IMatcher.hpp:19-24).  `matches` is a Python list that is APPENDED to and
    # never cleared; an empty side returns without touching it (CudaMatcher.cpp:35-37, :65).
    def match(self, query: Frame, train: Frame, matches: list, ratio_threshold=0.75):
        if len(query.descriptors) == 0 or len(train.descriptors) == 0:
            return
        q = np.ascontiguousarray(query.descriptors, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(train.descriptors, np.uint8).reshape(-1, 32)
        out = np.zeros(len(q), MATCH_DTYPE); n = C.c_int(0)
        check(lib().aslam_match_knn2(self.ctx.h, ptr(q), len(q), ptr(t), len(t), C.c_float(ratio_threshold), ptr(out), C.byref(n)))
        matches.extend(out[:n.value])

    # The legacy tracking loop's form (src/main.cpp:158-175): ratio test, then drop matches with an endpoint on a
    # dynamic object.  Appends like match(); returns main.cpp's filtered_count.
    def matchDynamic(self, query: Frame, train: Frame, detections, matches: list, ratio_threshold=0.75, half_open=True):
        if len(query.descriptors) == 0 or len(train.descriptors) == 0:
            return 0
        q = np.ascontiguousarray(query.descriptors, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(train.descriptors, np.uint8).reshape(-1, 32)
        qk = np.ascontiguousarray(query.keypoints, KP_DTYPE); tk = np.ascontiguousarray(train.keypoints, KP_DTYPE)
        b = None if detections is None or len(detections) == 0 else as_boxes(detections)
        out = np.zeros(len(q), MATCH_DTYPE); n = C.c_int(0); nf = C.c_int(0)
        check(lib().aslam_match_knn2_dynamic(self.ctx.h, ptr(q), ptr(qk), len(q), ptr(t), ptr(tk), len(t),
                                             C.c_float(ratio_threshold), ptr(b), 0 if b is None else len(b),
                                             1 if half_open else 0, ptr(out), C.byref(n), C.byref(nf)))
        matches.extend(out[:n.value])
        return nf.value

    # IMatcher::matchMultiple (IMatcher.hpp:27-37) as one batched launch
    def matchMultiple(self, query: Frame, candidates: List[Frame], all_matches: list, ratio_threshold=0.75):
        while len(all_matches) < len(candidates):
            all_matches.append([])
        del all_matches[len(candidates):]
        if len(query.descriptors) == 0 or not candidates:
            return
        q = np.ascontiguousarray(query.descriptors, np.uint8).reshape(-1, 32)
        sets = [np.ascontiguousarray(c.descriptors, np.uint8).reshape(-1, 32) for c in candidates]
        start = np.zeros(len(sets) + 1, np.int32)
        start[1:] = np.cumsum([len(s) for s in sets])
        t = np.concatenate(sets) if start[-1] else np.zeros((0, 32), np.uint8)
        out = np.zeros((len(sets), len(q)), MATCH_DTYPE); nout = np.zeros(len(sets), np.int32)
        check(lib().aslam_match_multi(self.ctx.h, ptr(q), len(q), ptr(t), ptr(start), len(sets),
                                      C.c_float(ratio_threshold), ptr(out), ptr(nout)))
        for i in range(len(sets)):
            all_matches[i].extend(out[i, :nout[i]])

    def knn2(self, q, t):
        """Raw neighbours (idx[nq,2], dist[nq,2]) -- the knnMatch(k=2) result before the ratio test."""
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
        idx = np.zeros((len(q), 2), np.int32); dist = np.zeros((len(q), 2), np.int32)
        check(lib().aslam_knn2_raw(self.ctx.h, ptr(q), len(q), ptr(t), len(t), ptr(idx), ptr(dist)))
        return idx, dist


# ----------------------------------------------------------------------------- loop detector
class LoopDB:
    """aslam_loopdb: keyframe descriptors resident in HBM (one shard)."""

    def __init__(self, ctx: Context, capacity_rows, max_keyframes=500):
        self.ctx = ctx
        h = C.c_void_p()
        check(lib().aslam_loopdb_create(ctx.h, int(capacity_rows), int(max_keyframes), C.byref(h)))
        self.h = h

    def add(self, desc, keyframe_id):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        check(lib().aslam_loopdb_add(self.h, ptr(d), len(d), int(keyframe_id)))

    def add_at(self, desc, keyframe_id, global_row_base):
        """A keyframe whose rows sit at global rows global_row_base.. (shards with scattered keyframes)."""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        check(lib().aslam_loopdb_add_at(self.h, ptr(d), len(d), int(keyframe_id), int(global_row_base)))

    def add_device(self, d_desc, n, keyframe_id):
        check(lib().aslam_loopdb_add_device(self.h, ptr(d_desc), int(n), int(keyframe_id)))

    def clear(self):
        check(lib().aslam_loopdb_clear(self.h))

    def save(self, path):
        check(lib().aslam_loopdb_save(self.h, str(path).encode()))

    def load(self, path):
        check(lib().aslam_loopdb_load(self.h, str(path).encode()))

    def set_row_base(self, base):
        check(lib().aslam_loopdb_set_row_base(self.h, int(base)))

    @property
    def num_keyframes(self):
        return lib().aslam_loopdb_num_keyframes(self.h)

    @property
    def num_rows(self):
        return lib().aslam_loopdb_num_rows(self.h)

    def keyframes(self):
        n = self.num_keyframes
        ids = np.zeros(n, np.uint64); rows = np.zeros(n, np.int32)
        check(lib().aslam_loopdb_keyframes(self.h, ptr(ids), ptr(rows)))
        return ids, rows

    def query_segmented(self, q, ratio=0.7):
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        good = np.zeros(max(self.num_keyframes, 1), np.int32)
        check(lib().aslam_loopdb_query_segmented(self.h, ptr(q), len(q), C.c_double(ratio), ptr(good)))
        return good[:self.num_keyframes]

    def query_segmented_device(self, d_q, nq, d_good, ratio=0.7):
        """Queries and counts stay on the device (d_good[num_keyframes] int32); asynchronous."""
        check(lib().aslam_loopdb_query_segmented_device(self.h, ptr(d_q), int(nq), C.c_double(ratio), ptr(d_good)))

    def match_keyframe(self, kf_index, q, ratio=0.7):
        """verifyGeometry's re-match (LoopClosure.cpp:116-130) against the resident rows of keyframe kf_index."""
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        out = np.zeros(max(len(q), 1), MATCH_DTYPE); n = C.c_int(0)
        check(lib().aslam_loopdb_match_keyframe(self.h, int(kf_index), ptr(q), len(q), C.c_double(ratio), ptr(out), C.byref(n)))
        return out[:n.value].copy()

    @property
    def capacity_rows(self):
        return lib().aslam_loopdb_capacity_rows(self.h)

    def query_global(self, q):
        """(row[nq,2] int64 global rows, dist[nq,2]); -1 where empty."""
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        top2 = np.zeros((len(q), 2), np.uint64)
        check(lib().aslam_loopdb_query_global(self.h, ptr(q), len(q), ptr(top2)))
        return unpack_top2(top2)

    def set_pipelining(self, on=True):
        """Consecutive fused queries may overlap (aslam_loopdb_set_pipelining; see the header for the caller's contract)."""
        check(lib().aslam_loopdb_set_pipelining(self.h, 1 if on else 0))

    def exchange_status(self, nqblocks=0, reset=False):
        """(status word, stamps[nqblocks, 4]) of the fused exchange: per query block the %globaltimer ns of the last push and
        of the last all-peers-seen, the sum of (seen - push) since the last reset, and the number of exchanges."""
        st = C.c_uint32(0)
        stamps = np.zeros((max(nqblocks, 1), 4), np.uint64)
        check(lib().aslam_loopdb_exchange_status(self.h, C.byref(st), ptr(stamps), int(nqblocks), 1 if reset else 0))
        return st.value, stamps[:nqblocks]

    def mean_exchange_wait_us(self, nqblocks):
        """Mean time per exchange this rank waited for its slowest peer (worst query block), since the last reset."""
        st, stamps = self.exchange_status(nqblocks)
        if st != 0:
            return -1.0
        n = stamps[:, 3].astype(np.float64)
        return float((stamps[:, 2].astype(np.float64) / np.maximum(n, 1)).max()) / 1e3

    # one process per GPU: peers' exchange buffers through CUDA IPC (aslam_loopdb_shard_export / _init)
    def shard_export(self, nranks, max_nq):
        h = np.zeros(lib().aslam_loopdb_shard_handle_bytes(), np.uint8)
        check(lib().aslam_loopdb_shard_export(self.h, int(nranks), int(max_nq), ptr(h)))
        return h

    def shard_init(self, rank, nranks, handles):
        handles = np.ascontiguousarray(handles, np.uint8).reshape(nranks, -1)
        check(lib().aslam_loopdb_shard_init(self.h, int(rank), int(nranks), ptr(handles)))

    def query_global_sharded_device(self, d_q, nq, d_out):
        check(lib().aslam_loopdb_query_global_sharded_device(self.h, ptr(d_q), int(nq), ptr(d_out)))

    def query_global_device(self, d_q, nq, d_top2):
        check(lib().aslam_loopdb_query_global_device(self.h, ptr(d_q), int(nq), ptr(d_top2)))

    def query_global_push_device(self, d_q, nq, d_peer_ptrs, nparts, my_part):
        check(lib().aslam_loopdb_query_global_push_device(self.h, ptr(d_q), int(nq), ptr(d_peer_ptrs), int(nparts), int(my_part)))

    def query_global_exchange_device(self, d_q, nq, d_peer_ptrs, nparts, my_part, round_, d_out):
        check(lib().aslam_loopdb_query_global_exchange_device(self.h, ptr(d_q), int(nq), ptr(d_peer_ptrs), int(nparts),
                                                              int(my_part), int(round_), ptr(d_out)))

    def close(self):
        if self.h:
            lib().aslam_loopdb_destroy(self.h)
            self.h = None


class LoopDBGroup:
    """aslam_loopdb_shard_group: ONE process, one shard per listed GPU (peer access between them, keyframe k on shard
    k % n).  Same calls and semantics as LoopDB; keyframes are indexed oldest first across the whole group."""

    def __init__(self, devices, capacity_rows_per_shard, max_keyframes=500):
        dev = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        check(lib().aslam_loopdb_shard_group_create(dev, len(devices), int(capacity_rows_per_shard), int(max_keyframes), C.byref(h)))
        self.h = h

    def add(self, desc, keyframe_id):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        check(lib().aslam_loopdb_shard_group_add(self.h, ptr(d), len(d), int(keyframe_id)))

    def clear(self):
        check(lib().aslam_loopdb_shard_group_clear(self.h))

    @property
    def num_shards(self):
        return lib().aslam_loopdb_shard_group_num_shards(self.h)

    @property
    def num_keyframes(self):
        return lib().aslam_loopdb_shard_group_num_keyframes(self.h)

    @property
    def num_rows(self):
        return lib().aslam_loopdb_shard_group_num_rows(self.h)

    def keyframes(self, with_shards=False):
        n = self.num_keyframes
        ids = np.zeros(n, np.uint64); rows = np.zeros(n, np.int32); shards = np.zeros(n, np.int32)
        check(lib().aslam_loopdb_shard_group_keyframes(self.h, ptr(ids), ptr(rows), ptr(shards)))
        return (ids, rows, shards) if with_shards else (ids, rows)

    def query_segmented(self, q, ratio=0.7):
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        good = np.zeros(max(self.num_keyframes, 1), np.int32)
        check(lib().aslam_loopdb_shard_group_query_segmented(self.h, ptr(q), len(q), C.c_double(ratio), ptr(good)))
        return good[:self.num_keyframes]

    def query_global(self, q, raw=False):
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        top2 = np.zeros((len(q), 2), np.uint64)
        check(lib().aslam_loopdb_shard_group_query_global(self.h, ptr(q), len(q), ptr(top2)))
        return top2 if raw else unpack_top2(top2)

    def read_result(self, shard, nq):
        top2 = np.zeros((nq, 2), np.uint64)
        check(lib().aslam_loopdb_shard_group_read_result(self.h, int(shard), int(nq), ptr(top2)))
        return top2

    def match_keyframe(self, kf_index, q, ratio=0.7):
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        out = np.zeros(max(len(q), 1), MATCH_DTYPE); n = C.c_int(0)
        check(lib().aslam_loopdb_shard_group_match_keyframe(self.h, int(kf_index), ptr(q), len(q), C.c_double(ratio), ptr(out), C.byref(n)))
        return out[:n.value].copy()

    def close(self):
        if self.h:
            lib().aslam_loopdb_shard_group_destroy(self.h)
            self.h = None


def unpack_top2(top2):
    top2 = np.asarray(top2, np.uint64)
    empty = top2 == np.uint64(0xFFFFFFFFFFFFFFFF)
    rows = (top2 & np.uint64((1 << 40) - 1)).astype(np.int64)
    dist = (top2 >> np.uint64(40)).astype(np.int64)
    rows[empty] = -1; dist[empty] = -1
    return rows, dist


class B200LoopDetector:
    """ILoopDetector over aslam_loopdb_* -- the descriptor side of LoopClosureDetector
    (src/legacy/LoopClosure.cpp:15-130).  detect() ranks keyframes exactly as findCandidates does
    and returns the best one that clears min_score/min_matches; the candidate re-match of verifyGeometry
    (:116-130) runs against the rows already resident in the database (no host copy of the keyframes, no
    upload of keyframe descriptors); the RANSAC check after it (:131-156) is host code outside this path.

    devices = a list of CUDA devices: the database is sharded over them inside this one process
    (aslam_loopdb_shard_group); None = one database on the context's device."""

    def __init__(self, ctx: Optional[Context] = None, min_frames_between=30, min_score=0.3, min_matches=30,
                 capacity_rows=None, max_keyframes=500, max_features=2000, devices=None):
        if capacity_rows is None:                       # the store grows on demand; this is only the first allocation
            capacity_rows = max_keyframes * (max_features + 64)
        if devices is not None and len(devices) > 1:
            self.ctx = None
            per_shard = -(-capacity_rows // len(devices))
            self.db = LoopDBGroup(devices, per_shard, max_keyframes)
        else:
            self.ctx = ctx or Context(devices[0] if devices else 0)
            self.db = LoopDB(self.ctx, capacity_rows, max_keyframes)
        self.min_frames_between = min_frames_between
        self.min_score = min_score
        self.min_matches = min_matches
        self._loops = 0

    def addKeyFrame(self, kf: KeyFrame):
        d = np.ascontiguousarray(kf.frame.descriptors, np.uint8).reshape(-1, 32)
        self.db.add(d, kf.id)

    def findCandidates(self, query: KeyFrame):
        """LoopClosure.cpp:72-114: [(keyframe index, score)], at most 5, best first."""
        q = np.ascontiguousarray(query.frame.descriptors, np.uint8).reshape(-1, 32)
        if len(q) == 0:
            return []
        good = self.db.query_segmented(q, 0.7)
        ids, rows = self.db.keyframes()
        out = []
        for i in range(len(ids)):
            if int(query.id) - int(ids[i]) < self.min_frames_between or rows[i] == 0:
                continue
            s = float(good[i]) / max(1, query.frame.numKeypoints())
            if s > 0.1:
                out.append((i, s))
        out.sort(key=lambda t: (-t[1], t[0]))
        return out[:5]

    def detect(self, query: KeyFrame) -> Optional[LoopCandidate]:
        if self.db.num_keyframes < self.min_frames_between:     # LoopClosure.cpp:34-36
            return None
        ids, _ = self.db.keyframes()
        q = np.ascontiguousarray(query.frame.descriptors, np.uint8).reshape(-1, 32)
        for i, score in self.findCandidates(query):
            if score < self.min_score:
                continue
            m = self.db.match_keyframe(i, q, 0.7)                # verifyGeometry's re-match (:120-128) on resident rows
            if len(m) < self.min_matches:
                continue
            self._loops += 1
            return LoopCandidate(int(query.id), int(ids[i]), score, m)
        return None

    def getLoopCount(self):
        return self._loops

    def setMinFramesBetween(self, n):
        self.min_frames_between = int(n)

    def setMinScore(self, s):
        self.min_score = float(s)

    def setMinMatches(self, n):
        self.min_matches = int(n)

    def close(self):
        self.db.close()


# ----------------------------------------------------------------------------- batched tracker
class OrbBatchTracker:
    """aslam_orb_track_*: extract a batch of equally sized frames and match each against its
    predecessor in one stream.  This is the high-throughput form of processFrame's
    extract -> match that the frames/s metric is measured on."""

    def __init__(self, width, height, max_features=2000, max_batch=256, ctx: Optional[Context] = None, blur_mode=2,
                 grid_cell=0, params: Optional[dict] = None):
        self.ctx = ctx or Context(0)
        p = _lib.OrbParams()
        lib().aslam_orb_default_params(C.byref(p))
        p.nfeatures = int(max_features); p.blur_mode = blur_mode; p.grid_cell = int(grid_cell)
        for name, value in (params or {}).items():      # any other cv::ORB::create argument (aslam_orb_params field)
            if name not in dict(_lib.OrbParams._fields_):
                raise ValueError("unknown ORB parameter %r" % name)
            setattr(p, name, value)
        h = C.c_void_p()
        check(lib().aslam_orb_create(self.ctx.h, C.byref(p), width, height, max_batch, C.byref(h)))
        self.h = h
        self.width, self.height, self.max_batch = width, height, max_batch
        self.capacity = lib().aslam_orb_capacity(h)
        self._n = 0

    def extract_async(self, images, masks=None, on_device=False):
        n = len(images)
        fs = self.width * self.height
        fn = lib().aslam_orb_extract_batch_device if on_device else lib().aslam_orb_extract_batch_async
        check(fn(self.h, ptr(images), n, fs, self.width, ptr(masks)))
        self._n = n

    def track_async(self, images, ratio=0.75, on_device=False):
        n = len(images)
        check(lib().aslam_orb_track_batch_async(self.h, ptr(images), n, self.width * self.height, self.width,
                                                C.c_float(ratio), 1 if on_device else 0))
        self._n = n

    def sync(self, kps=None, desc=None, counts=None, matches=None, match_counts=None, allow_overflow=True):
        """Copies into the given host arrays (numpy or pinned torch tensors); any may be None."""
        rc = lib().aslam_orb_track_sync(self.h, ptr(kps), ptr(desc), ptr(counts), ptr(matches), ptr(match_counts))
        return check(rc, allow_overflow=allow_overflow)

    def results(self, with_matches=False):
        n, cap = self._n, self.capacity
        kps = np.zeros((n, cap), KP_DTYPE); desc = np.zeros((n, cap, 32), np.uint8); cnt = np.zeros(n, np.int32)
        m = np.zeros((n, cap), MATCH_DTYPE) if with_matches else None
        mc = np.zeros(n, np.int32) if with_matches else None
        self.sync(kps, desc, cnt, m, mc)
        return kps, desc, cnt, m, mc

    def device_results(self):
        k, d, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().aslam_orb_device_results(self.h, C.byref(k), C.byref(d), C.byref(c)))
        return k.value, d.value, c.value

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        check(lib().aslam_orb_level_size(self.h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def read_level(self, frame, level, blurred=False):
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        check(lib().aslam_orb_read_level(self.h, frame, level, 1 if blurred else 0, ptr(out)))
        return out

    def read_corners(self, frame, level):
        w, h = self.level_size(level)
        cap = w * h // 4 + 16
        out = np.zeros(cap, np.uint32); n = C.c_int()
        check(lib().aslam_orb_read_corners(self.h, frame, level, ptr(out), cap, C.byref(n)))
        e = out[:n.value]
        return (e & 0xfff).astype(np.int32), ((e >> 12) & 0xfff).astype(np.int32), (e >> 24).astype(np.int32)

    def close(self):
        if self.h:
            lib().aslam_orb_destroy(self.h)
            self.h = None
