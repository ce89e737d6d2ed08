"""N > 1 on real GPUs (skipped on a single-GPU box; the exchange steps are covered over gloo in test_host_logic.py and the
fused kernel with one shard in test_gpu_match.py): the sharded loop query in both process models must equal the CPU oracle
on every rank / device -- full [nq, 2] keys, and the per-keyframe counts."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")


def test_sharded_loop_query_one_process_per_gpu():
    _two_gpus()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "multi_gpu_check.py"), "60000"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["modes_agree"] and d["second_neighbour_checked"]
    for mode in ("nccl", "push", "fused", "ipc", "ipc+pipelined"):
        e = d[mode]
        assert e["exchange"] == mode.split("+")[0] and e["ranks_identical"] and e["every_rank_equals_oracle_full_keys"], (mode, e)
        assert e["segmented_equals_oracle"] and e["flat_after_segmented_ok"], (mode, e)
    assert d["ipc"]["status"] == 0 and d["fused"]["status"] == 0


def test_shard_group_two_devices_in_one_process(ctx, oracle):
    """aslam_loopdb_shard_group over two GPUs from ONE host thread: round-robin keyframe placement under FIFO eviction,
    per-keyframe counts, flat top-2 (identical on both devices), candidate re-match on the owning shard."""
    _two_gpus()
    from aria_slam_b200.frontend import LoopDB, LoopDBGroup
    rs = np.random.RandomState(101)
    sizes = [rs.randint(150, 400) if i % 6 else 1 for i in range(40)]
    kfs = [rs.randint(0, 256, (n, 32), dtype=np.uint8) for n in sizes]
    grp = LoopDBGroup([0, 1], 2000, 15)
    one = LoopDB(ctx, 4000, 15)
    for i, k in enumerate(kfs):
        grp.add(k, i)
        one.add(k, i)
        if i in (9, 22, 39):
            live = kfs[max(0, i - 14):i + 1]
            flat = np.concatenate(live)
            q = np.concatenate([live[0][:40], live[len(live) // 2][:40], live[-1][:40], rs.randint(0, 256, (30, 32), dtype=np.uint8)])
            q[::3, 17] ^= 0x24
            start = np.concatenate([[0], np.cumsum([len(k2) for k2 in live])]).astype(np.int32)
            good = grp.query_segmented(q, 0.7)
            assert (good == oracle.loop_segment_counts(q, flat, start, 7, 10)).all() and (good == one.query_segmented(q, 0.7)).all()
            keys = grp.query_global(q, raw=True)
            assert (grp.read_result(0, len(q)) == keys).all() and (grp.read_result(1, len(q)) == keys).all()
            from aria_slam_b200.frontend import unpack_top2
            (r1, d1), (r2, d2) = unpack_top2(keys), one.query_global(q)
            assert (r1 == r2).all() and (d1 == d2).all()
            ridx, rdist = oracle.knn2(q, flat)
            assert (r1 - sum(len(k2) for k2 in kfs[:max(0, i - 14)]) == ridx).all() and (d1 == rdist).all()
            for kf in (0, len(live) // 2, len(live) - 1):
                assert (grp.match_keyframe(kf, q, 0.7) == one.match_keyframe(kf, q, 0.7)).all()
    ids, rows, shards = grp.keyframes(with_shards=True)
    assert list(ids) == list(range(25, 40)) and list(shards) == [i % 2 for i in range(25, 40)]
    grp.close(); one.close()


def test_sharded_loop_query_between_processes_without_torch():
    """adapters/shard_ipc_selftest: two processes forked by a CUDA-free parent, CUDA-IPC handles exchanged through pipes, the C
    ABI only (no torch, no NCCL, no CUDA runtime in the program); every rank's keys equal the unsharded database's."""
    _two_gpus()
    exe = os.path.join(ROOT, "aria_slam_b200", "adapters", "shard_ipc_selftest")
    r = subprocess.run([exe, "2", "120000", "2000", "6"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:]
