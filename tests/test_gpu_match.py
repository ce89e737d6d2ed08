"""GPU parity of Hamming kNN / ratio matching / loop database against the CPU oracle
(bit-exact: distances, indices, ratio survivors).  All calls go through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_desc(n, seed, entropy=256):
    rs = np.random.RandomState(seed)
    d = rs.randint(0, 256, (n, 32), dtype=np.uint8)
    if entropy < 256:           # low-entropy descriptors: few distinct bytes -> mass distance ties
        d = (d % entropy).astype(np.uint8)
    return d


@pytest.mark.parametrize("nq,nt,seed", [(1000, 1000, 1), (2000, 2000, 2), (513, 129, 3), (1, 2, 4), (7, 4097, 5),
                                        (4000, 4000, 6), (300, 1, 7), (64, 128, 8)])
def test_knn2_raw_matches_oracle(ctx, oracle, nq, nt, seed):
    from aria_slam_b200.frontend import B200Matcher
    q, t = _rand_desc(nq, seed), _rand_desc(nt, seed + 100)
    idx, dist = B200Matcher(ctx).knn2(q, t)
    ridx, rdist = oracle.knn2(q, t)
    assert (idx == ridx).all() and (dist == rdist).all()


def test_knn2_ties_go_to_lowest_index(ctx, oracle):
    from aria_slam_b200.frontend import B200Matcher
    t = _rand_desc(900, 11, entropy=2)           # bytes in {0,1}: distances collide constantly
    t[100:130] = t[5]                            # exact duplicates of row 5
    q = np.concatenate([t[5:6], t[700:800], _rand_desc(50, 12, entropy=2)])
    idx, dist = B200Matcher(ctx).knn2(q, t)
    ridx, rdist = oracle.knn2(q, t)
    assert (idx == ridx).all() and (dist == rdist).all()
    assert idx[0, 0] == 5 and idx[0, 1] == 100 and dist[0, 0] == 0 and dist[0, 1] == 0


@pytest.mark.parametrize("ratio,num,den", [(0.75, 3, 4), (0.7, None, None), (0.9, None, None), (1.0, 1, 1)])
def test_match_ratio_survivors(ctx, oracle, ratio, num, den):
    from aria_slam_b200.frontend import B200Matcher, Frame
    t = _rand_desc(1500, 21)
    q = t[::2].copy()
    rs = np.random.RandomState(22)
    for i in range(len(q)):                      # 0..60 flipped bits -> a spread of d0/d1 ratios
        bits = rs.choice(256, rs.randint(0, 61), replace=False)
        np.bitwise_xor.at(q[i], bits // 8, (1 << (bits % 8)).astype(np.uint8))
    m = []
    B200Matcher(ctx).match(Frame(descriptors=q.reshape(-1), keypoints=np.zeros(len(q))),
                           Frame(descriptors=t.reshape(-1), keypoints=np.zeros(len(t))), m, ratio)
    idx, dist = oracle.knn2(q, t)
    # CudaMatcher.cpp:60 in float arithmetic
    ok = dist[:, 0].astype(np.float32) < np.float32(ratio) * dist[:, 1].astype(np.float32)
    ref = [(i, idx[i, 0], float(dist[i, 0])) for i in np.nonzero(ok)[0]]
    got = [(int(x["query_idx"]), int(x["train_idx"]), float(x["distance"])) for x in m]
    assert got == ref
    if num:
        assert len(got) == len(oracle.ratio_matches(idx, dist, num, den))


def test_match_appends_and_empty_sides(ctx):
    from aria_slam_b200.frontend import B200Matcher, Frame
    mt = B200Matcher(ctx)
    q = Frame(descriptors=_rand_desc(10, 31).reshape(-1))
    sentinel = ["keep"]
    mt.match(q, Frame(), sentinel)                    # CudaMatcher.cpp:35-37: untouched
    mt.match(Frame(), q, sentinel)
    assert sentinel == ["keep"]
    one = Frame(descriptors=_rand_desc(1, 32).reshape(-1))
    mt.match(q, one, sentinel, 0.99)                  # nt < 2: knn.size() >= 2 never holds
    assert sentinel == ["keep"]
    mt.match(q, q, sentinel, 0.99)                    # appends, never clears (:65)
    assert sentinel[0] == "keep" and len(sentinel) == 11
    before = len(sentinel)
    mt.match(q, q, sentinel, 0.0)                     # implementation semantics of 0.0: nothing passes
    assert len(sentinel) == before


DYNAMIC = {0, 1, 2, 3, 5, 6, 7, 14, 15, 16}          # src/main.cpp:29-41


def _in_dynamic(x, y, boxes, half_open):
    """isInDynamicObject (src/main.cpp:43-50) on float32 values, with cv::Rect::contains (half-open) or
    core::Detection::contains (closed, include/core/Types.hpp:109-111)."""
    for x1, y1, x2, y2, _, cls in boxes:
        if int(cls) in DYNAMIC and x >= x1 and y >= y1 and ((x < x2 and y < y2) if half_open else (x <= x2 and y <= y2)):
            return True
    return False


@pytest.mark.parametrize("half_open", [True, False])
def test_match_dynamic_follows_the_legacy_tracking_loop(ctx, oracle, half_open):
    """src/main.cpp:158-175: knnMatch(k=2) + 0.75 ratio, then a match is kept only if neither its query keypoint nor
    its train keypoint lies in a dynamic-class detection; filtered_count counts the rest.  Real extracted frames, boxes
    whose edges pass exactly through keypoint coordinates (the half-open / closed difference) and a static class."""
    from aria_slam_b200 import synth
    from aria_slam_b200.frontend import OrbB200Extractor, B200Matcher, Frame
    a = synth.frame(640, 480, 51); b = synth.shifted(a, seed=52)
    ex = OrbB200Extractor(1000, ctx)
    fa, fb = Frame(), Frame()
    ex.extract(a, 640, 480, fa); ex.extract(b, 640, 480, fb)
    ex.close()
    _, boxes = synth.mask_and_boxes(640, 480, 53)
    boxes[1, 5] = 9                                       # traffic light: static, must not filter anything
    k = fb.keypoints[np.argsort(fb.keypoints["x"])[len(fb.keypoints) // 2]]
    boxes[2, :4] = (k["x"] - 40, k["y"] - 30, k["x"], k["y"])          # a keypoint exactly on the far corner
    k = fa.keypoints[7]
    boxes[3, :4] = (k["x"], k["y"], k["x"] + 25, k["y"] + 35)          # and one exactly on the near corner
    mt = B200Matcher(ctx)
    got = []
    nf = mt.matchDynamic(fb, fa, boxes, got, 0.75, half_open=half_open)
    idx, dist = oracle.knn2(fb.descriptors.reshape(-1, 32), fa.descriptors.reshape(-1, 32))
    survivors = oracle.ratio_matches(idx, dist, 3, 4)
    keep, dropped = [], 0
    for qi, ti, d in survivors.tolist():
        p1, p2 = fb.keypoints[qi], fa.keypoints[ti]
        if _in_dynamic(p1["x"], p1["y"], boxes, half_open) or _in_dynamic(p2["x"], p2["y"], boxes, half_open):
            dropped += 1
        else:
            keep.append((qi, ti, d))
    assert [(int(m["query_idx"]), int(m["train_idx"]), int(m["distance"])) for m in got] == keep
    assert nf == dropped and dropped > 20 and len(keep) > 100
    # no boxes = IMatcher::match
    plain, same = [], []
    mt.match(fb, fa, plain, 0.75)
    assert mt.matchDynamic(fb, fa, None, same, 0.75) == 0
    assert [tuple(x) for x in same] == [tuple(x) for x in plain] and len(plain) == len(keep) + dropped


def test_match_dynamic_half_open_and_closed_differ_only_on_the_far_edge(ctx):
    from aria_slam_b200.frontend import B200Matcher, Frame
    from aria_slam_b200._lib import KP_DTYPE
    d = _rand_desc(64, 77)
    kps = np.zeros(64, KP_DTYPE)
    kps["x"] = np.arange(64, dtype=np.float32) * 2.5; kps["y"] = 10.0
    q = Frame(keypoints=kps, descriptors=d.reshape(-1)); t = Frame(keypoints=kps.copy(), descriptors=d.reshape(-1).copy())
    box = np.array([[25.0, 10.0, 50.0, 10.0, 0.9, 0]], np.float32)            # zero-height box on the keypoint row
    mt = B200Matcher(ctx)
    closed, half = [], []
    nf_closed = mt.matchDynamic(q, t, box, closed, 0.99, half_open=False)
    nf_half = mt.matchDynamic(q, t, box, half, 0.99, half_open=True)
    on_closed = int(((kps["x"] >= 25) & (kps["x"] <= 50)).sum())
    assert nf_closed == on_closed == 11 and len(closed) == 64 - 11
    assert nf_half == 0 and len(half) == 64                                    # y < y2 never holds for a zero-height rect


def test_match_multiple_equals_looped_match(ctx):
    from aria_slam_b200.frontend import B200Matcher, Frame
    mt = B200Matcher(ctx)
    q = Frame(descriptors=_rand_desc(700, 41).reshape(-1))
    cands = [Frame(descriptors=_rand_desc(n, 42 + i).reshape(-1)) for i, n in enumerate([500, 0, 1, 2, 1300])]
    cands[0].descriptors[:32 * 50] = q.descriptors[:32 * 50]
    allm = []
    mt.matchMultiple(q, cands, allm, 0.8)
    assert len(allm) == len(cands)
    for c, got in zip(cands, allm):
        ref = []
        mt.match(q, c, ref, 0.8)
        assert [tuple(x) for x in got] == [tuple(x) for x in ref]
    assert len(allm[0]) >= 50 and allm[1] == [] and allm[2] == []


def _make_db(nkf, per, seed):
    rs = np.random.RandomState(seed)
    sizes = [per if i % 7 else max(2, per // 3) for i in range(nkf)]
    sizes[3] = 1                                           # a keyframe with a single descriptor
    return [rs.randint(0, 256, (n, 32), dtype=np.uint8) for n in sizes]


def test_loopdb_segmented_counts(ctx, oracle):
    from aria_slam_b200.frontend import LoopDB
    kfs = _make_db(40, 300, 51)
    db = LoopDB(ctx, sum(len(k) for k in kfs) + 10, 500)
    for i, k in enumerate(kfs):
        db.add(k, i)
    flat = np.concatenate(kfs)
    q = flat[np.random.RandomState(52).randint(0, len(flat), 400)].copy()
    q[::3, 5] ^= 0x3c
    good = db.query_segmented(q, 0.7)
    start = np.concatenate([[0], np.cumsum([len(k) for k in kfs])]).astype(np.int32)
    ref = oracle.loop_segment_counts(q, flat, start, 7, 10)
    assert (good == ref).all()
    # LoopClosure.cpp:92 uses the double literal 0.7: same survivors as 10*d0 < 7*d1
    idx, dist = oracle.knn2(q, kfs[5])
    assert ref[5] == int(((dist[:, 0].astype(np.float64)) < 0.7 * dist[:, 1].astype(np.float64)).sum())
    db.close()


def test_loopdb_fifo_eviction_and_slide(ctx, oracle):
    from aria_slam_b200.frontend import LoopDB
    kfs = _make_db(30, 100, 61)
    db = LoopDB(ctx, 1200, 8)                   # small capacity: forces eviction and the slide-back
    for i, k in enumerate(kfs):
        db.add(k, 1000 + i)
    ids, rows = db.keyframes()
    assert list(ids) == [1000 + i for i in range(22, 30)]         # LoopClosure.cpp:27-30
    assert list(rows) == [len(k) for k in kfs[22:]]
    q = np.concatenate([kfs[25][:20], kfs[10][:20]])
    good = db.query_segmented(q, 0.7)
    live = np.concatenate(kfs[22:])
    start = np.concatenate([[0], np.cumsum([len(k) for k in kfs[22:]])]).astype(np.int32)
    assert (good == oracle.loop_segment_counts(q, live, start, 7, 10)).all()
    db.close()


def test_loopdb_global_top2_and_sharded_merge(ctx, oracle):
    """Shards of one database queried separately and merged == the unsharded query == the oracle."""
    import torch
    from aria_slam_b200 import synth
    from aria_slam_b200.frontend import LoopDB, unpack_top2
    from aria_slam_b200._lib import lib, check, ptr
    db_rows = synth.descriptor_db(60_000, 5)
    q, _ = synth.loop_queries(db_rows, 500, 6)
    ridx, rdist = oracle.knn2(q, db_rows)
    whole = LoopDB(ctx, len(db_rows), 10)
    whole.add(db_rows, 0)
    rows, dist = whole.query_global(q)
    assert (rows == ridx).all() and (dist == rdist).all()
    whole.close()
    for nshard in (2, 3, 8):
        bounds = np.linspace(0, len(db_rows), nshard + 1).astype(np.int64)
        parts = torch.empty((nshard, len(q), 2), dtype=torch.int64, device="cuda")
        dq = torch.from_numpy(q).cuda()
        torch.cuda.synchronize()
        shards = []
        for s in range(nshard):
            sh = LoopDB(ctx, int(bounds[s + 1] - bounds[s]) + 1, 10)
            sh.add(db_rows[bounds[s]:bounds[s + 1]], s)
            sh.set_row_base(int(bounds[s]))
            sh.query_global_device(dq, len(q), parts[s])
            shards.append(sh)
        out = torch.empty((len(q), 2), dtype=torch.int64, device="cuda")
        check(lib().aslam_loop_merge_top2_device(ctx.h, ptr(parts), nshard, len(q), ptr(out)))
        ctx.sync()
        mrows, mdist = unpack_top2(out.cpu().numpy().view(np.uint64))
        assert (mrows == ridx).all() and (mdist == rdist).all()
        for sh in shards:
            sh.close()


def test_loop_detector_candidates_follow_reference_ranking(ctx, oracle):
    from aria_slam_b200.frontend import B200LoopDetector, KeyFrame, Frame
    rs = np.random.RandomState(71)
    det = B200LoopDetector(ctx, min_frames_between=5, min_score=0.3, min_matches=10, capacity_rows=40 * 400)
    kfs = []
    for i in range(30):
        d = rs.randint(0, 256, (300, 32), dtype=np.uint8)
        kfs.append(d)
        det.addKeyFrame(KeyFrame(id=i, frame=Frame(descriptors=d.reshape(-1), keypoints=np.zeros(300))))
    qd = np.concatenate([kfs[4][:150], kfs[9][:60], kfs[28][:90]])      # 28 is inside the gap
    query = KeyFrame(id=31, frame=Frame(descriptors=qd.reshape(-1), keypoints=np.zeros(len(qd))))
    cands = det.findCandidates(query)
    flat = np.concatenate(kfs)
    start = (np.arange(31) * 300).astype(np.int32)
    good = oracle.loop_segment_counts(qd, flat, start, 7, 10)
    ref = oracle.loop_candidates(good, len(qd), list(range(30)), 31, 5)
    assert [c[0] for c in cands] == [c[0] for c in ref] == [4, 9]
    assert all(abs(a[1] - b[1]) == 0 for a, b in zip(cands, ref))
    hit = det.detect(query)
    assert hit is not None and hit.match_id == 4 and hit.query_id == 31 and len(hit.matches) >= 150
    assert det.getLoopCount() == 1


def test_loopdb_repeated_flat_queries_reuse_state_correctly(ctx, oracle):
    """The flat query keeps its row-range table and leaves the merged result slots preset between calls; a
    segmented query, a different query size or a database change in between must not leak into the next answer."""
    from aria_slam_b200 import synth
    from aria_slam_b200.frontend import LoopDB
    rows = synth.descriptor_db(40_000, 9)
    q1, _ = synth.loop_queries(rows, 300, 10)
    q2, _ = synth.loop_queries(rows, 777, 11)
    db = LoopDB(ctx, 60_000, 64)
    r0, d0 = db.query_global(q1)                                   # empty database: no neighbours at all
    assert (r0 == -1).all() and (d0 == -1).all()
    for k in range(20):
        db.add(rows[k * 2000:(k + 1) * 2000], k)
    ref1 = oracle.knn2(q1, rows)
    ref2 = oracle.knn2(q2, rows)
    for q, ref in ((q1, ref1), (q1, ref1), (q2, ref2), (q1, ref1)):
        r, d = db.query_global(q)
        assert (r == ref[0]).all() and (d == ref[1]).all()
    good = db.query_segmented(q1, 0.7)                             # overwrites the shared scratch
    start = np.arange(0, 40_001, 2000).astype(np.int32)
    assert (good == oracle.loop_segment_counts(q1, rows, start, 7, 10)).all()
    r, d = db.query_global(q1)
    assert (r == ref1[0]).all() and (d == ref1[1]).all()
    extra = synth.descriptor_db(2000, 12)
    db.add(extra, 20)                                              # the live region grows: the table is rebuilt
    allrows = np.concatenate([rows, extra])
    ref3 = oracle.knn2(q2, allrows)
    r, d = db.query_global(q2)
    assert (r == ref3[0]).all() and (d == ref3[1]).all()
    db.close()


def test_loopdb_save_load_round_trip(ctx, oracle, tmp_path):
    """The keyframe store written to a file and read into a fresh database answers every query identically
    (segmented counts, flat top-2, keyframe table, row base); a bad or oversized file is an error."""
    from aria_slam_b200 import _lib
    from aria_slam_b200.frontend import LoopDB
    rs = np.random.RandomState(21)
    kfs = [rs.randint(0, 256, (n, 32), dtype=np.uint8) for n in (300, 0, 1200, 777, 64)]
    q = np.concatenate([kfs[2][:100], rs.randint(0, 256, (150, 32), dtype=np.uint8)])
    a = LoopDB(ctx, 4000, 8)
    for i, k in enumerate(kfs):
        a.add(k, 100 + i)
    a.set_row_base(5000)
    path = tmp_path / "loop.db"
    a.save(path)
    b = LoopDB(ctx, 4000, 8)
    b.add(kfs[0], 1)                                   # replaced by load
    b.load(path)
    assert b.num_keyframes == 5 and b.num_rows == sum(len(k) for k in kfs)
    ia, ra = a.keyframes(); ib, rb = b.keyframes()
    assert (ia == ib).all() and (ra == rb).all()
    assert (a.query_segmented(q, 0.7) == b.query_segmented(q, 0.7)).all()
    (r1, d1), (r2, d2) = a.query_global(q), b.query_global(q)
    assert (r1 == r2).all() and (d1 == d2).all() and r1[:100, 0].min() >= 5000
    grown = LoopDB(ctx, 1000, 8)
    grown.load(path)                                   # more rows than its first allocation: the store grows
    assert grown.num_rows == b.num_rows and (grown.query_segmented(q, 0.7) == a.query_segmented(q, 0.7)).all()
    grown.close()
    small = LoopDB(ctx, 4000, 3)
    with pytest.raises(_lib.AslamError):
        small.load(path)                               # more keyframes than this database's FIFO cap
    bad = tmp_path / "bad.db"
    bad.write_bytes(b"not a database")
    with pytest.raises(_lib.AslamError):
        b.load(bad)
    assert b.num_keyframes == 5                        # a failed load leaves the database as it was
    for d in (a, b, small):
        d.close()


def test_match_keyframe_on_resident_rows_equals_the_oracle(ctx, oracle):
    """verifyGeometry's re-match (LoopClosure.cpp:116-130): knnMatch(query, keyframe) + d0 < 0.7*d1 in double against the
    rows already stored in the database, for keyframes before and after an eviction + slide."""
    from aria_slam_b200.frontend import LoopDB
    kfs = _make_db(12, 250, 81)
    db = LoopDB(ctx, 1500, 6)
    for i, k in enumerate(kfs):
        db.add(k, i)
    live = kfs[6:]
    rs = np.random.RandomState(82)
    q = np.concatenate([live[2][:120], live[4][:50], rs.randint(0, 256, (80, 32), dtype=np.uint8)])
    q[::4, 9] ^= 0x81
    for i, k in enumerate(live):
        m = db.match_keyframe(i, q, 0.7)
        if len(k) < 2:
            assert len(m) == 0
            continue
        idx, dist = oracle.knn2(q, k)
        ref = oracle.ratio_matches(idx, dist, 7, 10)
        got = np.stack([m["query_idx"], m["train_idx"], m["distance"].astype(np.int32)], 1)
        assert got.shape == ref.shape and (got == ref).all(), "keyframe %d" % i
    assert len(db.match_keyframe(2, q[:0], 0.7)) == 0
    db.close()


def test_fused_scan_exchange_with_one_shard_equals_the_plain_query(ctx, oracle):
    """The fused scan + exchange kernel with nparts = 1 (this GPU is its only peer) goes through every step of the
    protocol -- merged slots, push, flag, acquire, merge, re-arm -- and must give the plain flat query's keys, round
    after round, also after a segmented query reused the scratch and for a ragged number of query blocks."""
    import torch
    from aria_slam_b200 import synth
    from aria_slam_b200.frontend import LoopDB
    from aria_slam_b200._lib import lib
    from aria_slam_b200.sharded import pack_top2
    rows = synth.descriptor_db(30_000, 31)
    db = LoopDB(ctx, 40_000, 64)
    for k in range(15):
        db.add(rows[k * 2000:(k + 1) * 2000], k)
    db.set_row_base(7_000_000)
    for nq in (1300, 2000, 7):
        q, _ = synth.loop_queries(rows, nq, 32 + nq)
        ridx, rdist = oracle.knn2(q, rows)
        ref = pack_top2(np.where(ridx >= 0, ridx.astype(np.int64) + 7_000_000, -1), rdist)
        words = lib().aslam_loop_exchange_buffer_bytes(1, nq) // 8
        buf = torch.zeros(words, dtype=torch.int64, device="cuda")
        ptrs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device="cuda")
        dq = torch.from_numpy(q).cuda()
        out = torch.zeros((nq, 2), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        db.exchange_status(0, reset=True)
        for rnd in (1, 2, 3):
            out.zero_(); torch.cuda.synchronize()
            db.query_global_exchange_device(dq, nq, ptrs, 1, 0, rnd, out)
            ctx.sync()
            assert (out.cpu().numpy().view(np.uint64) == ref).all(), (nq, rnd)
            if rnd == 2:
                db.query_segmented(q[:50], 0.7)            # overwrites the armed slots and the row-range table
        st, stamps = db.exchange_status((nq + 511) // 512)
        assert st == 0 and (stamps[:, 1] >= stamps[:, 0]).all() and (stamps[:, 0] > 0).all() and (stamps[:, 3] == 3).all()
    db.close()


def test_pipelined_fused_queries_give_the_same_keys(ctx, oracle):
    """aslam_loopdb_set_pipelining: fused queries launched back to back (programmatic dependent launch: a query's scan may start
    while its predecessor's exchange tail runs) must leave exactly the keys of unpipelined queries -- also when the query
    buffer alternates (which ends the overlap by itself) and when a keyframe is added in between."""
    import torch
    from aria_slam_b200 import synth
    from aria_slam_b200.frontend import LoopDB
    from aria_slam_b200._lib import lib
    from aria_slam_b200.sharded import pack_top2
    rows = synth.descriptor_db(64_000, 41)
    db = LoopDB(ctx, 70_000, 64)
    for k in range(31):
        db.add(rows[k * 2000:(k + 1) * 2000], k)
    db.set_pipelining(True)
    nq = 2000
    qa, _ = synth.loop_queries(rows[:62_000], nq, 42)
    qb, _ = synth.loop_queries(rows[:62_000], nq, 43)
    refs = {}
    for name, q in (("a", qa), ("b", qb)):
        ridx, rdist = oracle.knn2(q, rows[:62_000])
        refs[name] = pack_top2(ridx, rdist)
    words = lib().aslam_loop_exchange_buffer_bytes(1, nq) // 8
    buf = torch.zeros(words, dtype=torch.int64, device="cuda")
    ptrs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device="cuda")
    da, dbq = torch.from_numpy(qa).cuda(), torch.from_numpy(qb).cuda()
    outs = [torch.zeros((nq, 2), dtype=torch.int64, device="cuda") for _ in range(12)]
    torch.cuda.synchronize()
    rnd = 0
    order = ["a"] * 6 + ["b", "a", "b", "b", "a", "a"]            # six pipelined repeats, then alternating buffers
    for i, name in enumerate(order):
        rnd += 1
        db.query_global_exchange_device(da if name == "a" else dbq, nq, ptrs, 1, 0, rnd, outs[i])
    ctx.sync()
    for i, name in enumerate(order):
        assert (outs[i].cpu().numpy().view(np.uint64) == refs[name]).all(), (i, name)
    db.add(rows[62_000:64_000], 31)                                # the database changes: the next query must see it
    ridx, rdist = oracle.knn2(qa, rows)
    for i in range(3):
        rnd += 1
        db.query_global_exchange_device(da, nq, ptrs, 1, 0, rnd, outs[i])
    ctx.sync()
    for i in range(3):
        assert (outs[i].cpu().numpy().view(np.uint64) == pack_top2(ridx, rdist)).all()
    assert db.exchange_status(0)[0] == 0
    db.close()


def test_shard_group_on_one_device_follows_the_single_database(ctx, oracle):
    """aslam_loopdb_shard_group with a single device: FIFO eviction, per-keyframe counts, flat top-2 with global rows
    that keep counting from creation, candidate re-match -- all equal to the plain aslam_loopdb and the oracle."""
    from aria_slam_b200.frontend import LoopDB, LoopDBGroup
    kfs = _make_db(25, 180, 91)
    grp = LoopDBGroup([ctx.device], 800, 9)
    one = LoopDB(ctx, 800, 9)
    for i, k in enumerate(kfs):
        grp.add(k, 500 + i)
        one.add(k, 500 + i)
    live = kfs[16:]
    ids, rows, shards = grp.keyframes(with_shards=True)
    assert list(ids) == [500 + i for i in range(16, 25)] and list(rows) == [len(k) for k in live] and (shards == 0).all()
    assert grp.num_shards == 1 and grp.num_keyframes == 9 and grp.num_rows == sum(len(k) for k in live)
    q = np.concatenate([live[1][:60], live[7][:60], kfs[2][:30]])
    q[::5, 3] ^= 0x18
    flat = np.concatenate(live)
    start = np.concatenate([[0], np.cumsum([len(k) for k in live])]).astype(np.int32)
    good = grp.query_segmented(q, 0.7)
    assert (good == one.query_segmented(q, 0.7)).all() and (good == oracle.loop_segment_counts(q, flat, start, 7, 10)).all()
    (r1, d1), (r2, d2) = grp.query_global(q), one.query_global(q)
    assert (r1 == r2).all() and (d1 == d2).all()
    ridx, rdist = oracle.knn2(q, flat)
    assert (r1 - sum(len(k) for k in kfs[:16]) == ridx).all() and (d1 == rdist).all()
    m1, m2 = grp.match_keyframe(7, q, 0.7), one.match_keyframe(7, q, 0.7)
    assert len(m1) >= 50 and (m1 == m2).all()
    grp.clear()
    assert grp.num_keyframes == 0 and len(grp.query_segmented(q, 0.7)) == 0
    grp.close(); one.close()


def test_round_robin_shards_with_scattered_keyframes_merge_to_the_unsharded_answer(ctx, oracle):
    """Keyframes dealt round-robin to 3 shards (aslam_loopdb_add_at: a shard's keyframes are scattered over the global row
    space, as the single-process shard group places them), with evictions: every shard translates its rows through its
    keyframe table, the merged flat top-2 equals the unsharded database's and the oracle's, including ties between
    duplicate rows that live on different shards."""
    import torch
    from aria_slam_b200.frontend import LoopDB, unpack_top2
    from aria_slam_b200._lib import lib, check, ptr
    rs = np.random.RandomState(111)
    kfs = [rs.randint(0, 256, (rs.randint(40, 700), 32), dtype=np.uint8) for _ in range(23)]
    kfs[11][:30] = kfs[4][:30]                      # duplicate rows on different shards: (distance, global row) ties
    kfs[17][5:25] = kfs[4][5:25]
    nshard = 3
    shards = [LoopDB(ctx, 2000, 64) for _ in range(nshard)]
    whole = LoopDB(ctx, 2000, 64)
    base = 0
    for i, k in enumerate(kfs):
        shards[i % nshard].add_at(k, i, base)
        whole.add(k, i)
        base += len(k)
    flat = np.concatenate(kfs)
    q = np.concatenate([kfs[4][:60], kfs[20][:50], rs.randint(0, 256, (90, 32), dtype=np.uint8)])
    q[::7, 2] ^= 0x41
    ridx, rdist = oracle.knn2(q, flat)
    parts = torch.empty((nshard, len(q), 2), dtype=torch.int64, device="cuda")
    dq = torch.from_numpy(q).cuda()
    torch.cuda.synchronize()
    for s_, sh in enumerate(shards):
        sh.query_global_device(dq, len(q), parts[s_])
    out = torch.empty((len(q), 2), dtype=torch.int64, device="cuda")
    check(lib().aslam_loop_merge_top2_device(ctx.h, ptr(parts), nshard, len(q), ptr(out)))
    ctx.sync()
    mrows, mdist = unpack_top2(out.cpu().numpy().view(np.uint64))
    wrows, wdist = whole.query_global(q)
    assert (mrows == ridx).all() and (mdist == rdist).all() and (wrows == ridx).all() and (wdist == rdist).all()
    # per-keyframe counts of the shards, put back in global keyframe order
    start = np.concatenate([[0], np.cumsum([len(k) for k in kfs])]).astype(np.int32)
    ref = oracle.loop_segment_counts(q, flat, start, 7, 10)
    got = np.zeros(len(kfs), np.int32)
    for s_, sh in enumerate(shards):
        got[s_::nshard] = sh.query_segmented(q, 0.7)
    assert (got == ref).all()
    from aria_slam_b200 import _lib
    with pytest.raises(_lib.AslamError):
        whole.add_at(kfs[0], 99, 10**6)             # a contiguous database does not take scattered keyframes
    with pytest.raises(_lib.AslamError):
        shards[0].add_at(kfs[0], 99, 5)             # global rows must increase within a shard
    for d in shards + [whole]:
        d.close()
