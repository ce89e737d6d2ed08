"""BASELINE.json's full sizes (C2: 256 frames of 752x480 with 2000 keypoints; C5: 2000 queries against a
1 M-row database), checked through properties that do not need the oracle at full size -- determinism,
invariance to how the batch is cut into frame streams / how the database is cut into shards -- plus
the oracle on a sample."""
import zlib

import numpy as np
import pytest

from conftest import kp_keys

pytestmark = pytest.mark.gpu


def _digest(kps, desc, cnt, m=None, mc=None):
    h = 0
    for i in range(len(cnt)):
        h = zlib.crc32(kps[i, :cnt[i]].tobytes(), h)
        h = zlib.crc32(desc[i, :cnt[i]].tobytes(), h)
        if m is not None:
            h = zlib.crc32(m[i, :mc[i]].tobytes(), h)
    return h


def test_c2_full_batch_is_deterministic_split_invariant_and_matches_the_oracle_on_a_sample(ctx, oracle):
    from aria_slam_b200 import synth
    from aria_slam_b200.frontend import OrbBatchTracker
    B, W, H, N = 256, 752, 480, 2000
    frames = synth.frame_stream(B, W, H, 1000)
    tr = OrbBatchTracker(W, H, N, B, ctx)
    tr.track_async(frames, 0.75)
    kps, desc, cnt, m, mc = tr.results(with_matches=True)
    first = _digest(kps, desc, cnt, m, mc)
    assert (cnt == N).all()                                    # the synthetic frames saturate every level quota
    # determinism: a second handle over the same frames gives byte-identical results (a checksum of checksums)
    tr2 = OrbBatchTracker(W, H, N, B, ctx)
    tr2.track_async(frames, 0.75)
    assert _digest(*tr2.results(with_matches=True)) == first
    tr2.close()
    # the bench cuts the batch into 8 frame streams: per-frame results must not depend on the cut; matches
    # differ only for the first frame of a stream (it has no predecessor there)
    S = 8
    sub = OrbBatchTracker(W, H, N, B // S, ctx)
    for s in range(S):
        lo = s * (B // S)
        sub2 = sub if s == 0 else OrbBatchTracker(W, H, N, B // S, ctx)
        sub2.track_async(frames[lo:lo + B // S], 0.75)
        k2, d2, c2, m2, mc2 = sub2.results(with_matches=True)
        assert (c2 == cnt[lo:lo + B // S]).all()
        for i in range(B // S):
            assert (k2[i, :c2[i]] == kps[lo + i, :c2[i]]).all() and (d2[i, :c2[i]] == desc[lo + i, :c2[i]]).all()
            if i:
                assert mc2[i] == mc[lo + i] and (m2[i, :mc2[i]] == m[lo + i, :mc2[i]]).all()
        assert mc2[0] == 0
        if s:
            sub2.close()
    sub.close()
    # the oracle on a sample of frames and matched pairs
    for i in (0, 97, 255):
        ko, do = oracle.orb(frames[i], N)
        assert sorted(kp_keys(kps[i, :cnt[i]])) == sorted(kp_keys(ko))
        og, oo = np.lexsort((kps[i, :N]["x"], kps[i, :N]["y"], kps[i, :N]["octave"])), np.lexsort((ko["x"], ko["y"], ko["octave"]))
        assert (desc[i, :N][og] == do[oo]).all(1).mean() >= 0.99      # the gate of north_star; 1.0 on every case so far
    for i in (1, 128, 255):
        idx, dist = oracle.knn2(desc[i - 1, :cnt[i - 1]], desc[i, :cnt[i]])      # query = previous frame (main.cpp:158-161)
        ref = oracle.ratio_matches(idx, dist, 3, 4)
        got = np.stack([m[i, :mc[i]]["query_idx"], m[i, :mc[i]]["train_idx"], m[i, :mc[i]]["distance"].astype(np.int32)], 1)
        assert got.shape == ref.shape and (got == ref).all()
    assert np.median(mc[1::2]) > 1000                          # odd frames are shifted copies of their predecessors
    tr.close()


def test_c5_million_row_database_shard_invariance_and_sampled_oracle(ctx, oracle):
    import torch
    from aria_slam_b200 import synth
    from aria_slam_b200.frontend import LoopDB, unpack_top2
    from aria_slam_b200._lib import lib, check, ptr
    rows_total, nq = 1_000_000, 2000
    db = synth.descriptor_db(rows_total, 5)
    q, src = synth.loop_queries(db, nq, 6)
    whole = LoopDB(ctx, rows_total, 500)
    for k in range(500):
        whole.add(db[k * 2000:(k + 1) * 2000], k)
    rows, dist = whole.query_global(q)
    whole.close()
    assert (rows[:, 0] == src).all()                           # the planted neighbours (<= 40 flipped bits) win
    assert (dist[:, 0] <= dist[:, 1]).all() and (dist[:, 0] <= 40).all()
    # a sample of queries against the oracle over the whole database
    pick = np.arange(0, nq, 125)
    ridx, rdist = oracle.knn2(q[pick], db)
    assert (rows[pick] == ridx).all() and (dist[pick] == rdist).all()
    # 8 keyframe-aligned shards (the 8-GPU layout) scanned one after another and merged == the unsharded answer
    from aria_slam_b200.sharded import plan_shards
    nshard = 8
    parts = torch.empty((nshard, nq, 2), dtype=torch.int64, device="cuda")
    dq = torch.from_numpy(q).cuda()
    torch.cuda.synchronize()
    for s, (lo, hi) in enumerate(plan_shards(500, nshard)):
        sh = LoopDB(ctx, (hi - lo) * 2000, hi - lo)
        for k in range(lo, hi):
            sh.add(db[k * 2000:(k + 1) * 2000], k)
        sh.set_row_base(lo * 2000)
        sh.query_global_device(dq, nq, parts[s])
        ctx.sync()
        sh.close()
    out = torch.empty((nq, 2), dtype=torch.int64, device="cuda")
    check(lib().aslam_loop_merge_top2_device(ctx.h, ptr(parts), nshard, nq, ptr(out)))
    ctx.sync()
    mrows, mdist = unpack_top2(out.cpu().numpy().view(np.uint64))
    assert (mrows == rows).all() and (mdist == dist).all()
