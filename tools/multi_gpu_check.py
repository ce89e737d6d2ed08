"""Run under torchrun on >= 2 GPUs: the push (fused finish + all-gather over peer memory) and NCCL
exchanges of the sharded loop query must both equal the CPU oracle, and are timed side by side."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from aria_slam_b200 import synth
from aria_slam_b200._lib import Context
from aria_slam_b200.frontend import unpack_top2
from aria_slam_b200.sharded import ShardedLoopQuery

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
db = synth.descriptor_db(rows, 5)
q, src = synth.loop_queries(db, 2000, 6)
res = {}
with torch.cuda.stream(stream):
    ctx = Context(local, stream.cuda_stream)
    dq = torch.from_numpy(q).cuda()
    for mode in ("nccl", "push", "fused"):
        lq = ShardedLoopQuery(ctx, db, rank, world, keyframe_rows=2000, exchange=mode)
        for _ in range(3):
            out = lq.query(dq, len(q))
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(20):
            out = lq.query(dq, len(q))
        e1.record(stream)
        torch.cuda.synchronize(); dist.barrier()
        r, d = unpack_top2(out.cpu().numpy().view(np.uint64))
        res[mode] = {"ms": e0.elapsed_time(e1) / 20, "best_is_source": bool((r[:, 0] == src).all()), "exchange": lq.exchange,
                     "keys": out.cpu().numpy().copy()}
same = bool((res["nccl"]["keys"] == res["push"]["keys"]).all() and (res["nccl"]["keys"] == res["fused"]["keys"]).all())
if rank == 0:
    if rows <= 300_000:
        from oracle import oracle as O
        ridx, rdist = O.knn2(q, db)
        r, d = unpack_top2(res["push"]["keys"].view(np.uint64))
        res["equals_oracle"] = bool((r == ridx).all() and (d == rdist).all())
    print(json.dumps({"world": world, "rows": rows, "push_equals_nccl": same,
                      **{k: {kk: vv for kk, vv in v.items() if kk != "keys"} if isinstance(v, dict) else v for k, v in res.items()}}))
dist.destroy_process_group()
