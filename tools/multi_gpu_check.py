"""Run under torchrun on >= 2 GPUs: every exchange of the sharded loop query (ipc = fused scan + exchange kernel over
CUDA-IPC buffers, fused = the same over symmetric memory, push, nccl) must give, on EVERY rank, the full [nq, 2] keys
(both neighbours, both distances) of the CPU oracle; the sharded per-keyframe query must give the oracle's counts.
Modes are timed side by side; the fused kernel's %globaltimer stamps give the time each rank waited for its slowest
peer.  Prints one JSON line (rank 0)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from aria_slam_b200 import synth
from aria_slam_b200._lib import Context
from aria_slam_b200.frontend import unpack_top2
from aria_slam_b200.sharded import ShardedLoopQuery, keyframe_starts, pack_top2

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["nccl", "push", "fused", "ipc", "ipc+pipelined"]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
NQ, KF = 2000, 2000
db = synth.descriptor_db(rows, 5)
q, src = synth.loop_queries(db, NQ, 6)
res = {}
ref_keys = None
if rows <= 300_000:                       # every rank checks its own copy against the oracle
    from oracle import oracle as O
    ridx, rdist = O.knn2(q, db)
    ref_keys = pack_top2(ridx, rdist)
    ref_good = O.loop_segment_counts(q, db, keyframe_starts(rows, KF).astype(np.int32), 7, 10)


def all_ranks(flag):
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


with torch.cuda.stream(stream):
    ctx = Context(local, stream.cuda_stream)
    dq = torch.from_numpy(q).cuda()
    for mode in modes:
        lq = ShardedLoopQuery(ctx, db, rank, world, keyframe_rows=KF, exchange=mode.split("+")[0], pipelined=mode.endswith("+pipelined"))
        for _ in range(3):
            out = lq.query(dq, NQ)
        torch.cuda.synchronize(); dist.barrier()
        if lq.exchange in ("ipc", "fused"):
            lq.db.exchange_status(0, reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            out = lq.query(dq, NQ)
        e1.record(stream)
        torch.cuda.synchronize(); dist.barrier()
        keys = out.cpu().numpy().view(np.uint64).copy()
        gathered = torch.empty((world,) + tuple(out.shape), dtype=out.dtype, device="cuda")
        dist.all_gather_into_tensor(gathered.view(-1, 2), out.contiguous())
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        r, d = unpack_top2(keys)
        ent = {"ms": float(ms.item()), "exchange": lq.exchange, "best_is_source": all_ranks((r[:, 0] == src).all()),
               "ranks_identical": bool((gathered == gathered[0]).all().item())}
        if ref_keys is not None:
            ent["every_rank_equals_oracle_full_keys"] = all_ranks((keys == ref_keys).all())
        if lq.exchange in ("ipc", "fused"):
            st = lq.db.exchange_status(0)[0]
            wait_us = lq.db.mean_exchange_wait_us((NQ + 511) // 512)
            w = torch.tensor([wait_us], dtype=torch.float64, device="cuda")
            allw = torch.empty(world, dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(allw, w)
            ent["status"] = st
            ent["mean_wait_for_slowest_peer_us_per_rank"] = [round(x, 1) for x in allw.cpu().tolist()]
        # the per-keyframe (reference-semantics) query, sharded: counts all-gathered
        for _ in range(2):
            good = lq.query_segmented(dq, NQ)
        torch.cuda.synchronize(); dist.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(steps):
            good = lq.query_segmented(dq, NQ)
        s1.record(stream)
        torch.cuda.synchronize(); dist.barrier()
        ent["segmented_ms"] = s0.elapsed_time(s1) / steps
        if ref_keys is not None:
            ent["segmented_equals_oracle"] = all_ranks((good.cpu().numpy() == ref_good).all())
        out2 = lq.query(dq, NQ)                     # the flat query after a segmented one (shared scratch) is still right
        torch.cuda.synchronize()
        ent["flat_after_segmented_ok"] = all_ranks((out2.cpu().numpy().view(np.uint64) == keys).all())
        res[mode] = ent
        res[mode + "_keys"] = keys
key_sets = [res[m + "_keys"] for m in modes]
same = all((k == key_sets[0]).all() for k in key_sets)
if rank == 0:
    out = {"world": world, "rows": rows, "queries": NQ, "modes_agree": bool(same),
           "second_neighbour_checked": ref_keys is not None}
    out.update({m: res[m] for m in modes})
    print(json.dumps(out))
dist.destroy_process_group()
