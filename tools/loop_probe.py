"""Loop-query scan timing on one GPU for a database of `rows` rows (the per-GPU shard of an N-GPU run)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aria_slam_b200._lib import Context
from aria_slam_b200.frontend import LoopDB
from aria_slam_b200 import synth

ctx = Context(0)
out = {}
for rows in [int(a) for a in sys.argv[1:]] or [1_000_000, 125_000]:
    db = synth.descriptor_db(rows, 5)
    q, src = synth.loop_queries(db, 2000, 6)
    ldb = LoopDB(ctx, rows, 600)
    for k in range(0, rows, 2000):
        ldb.add(db[k:k + 2000], k // 2000)
    dq = torch.from_numpy(q).cuda(); top = torch.empty((2000, 2), dtype=torch.int64, device="cuda")
    for _ in range(3):
        ldb.query_global_device(dq, 2000, top)
    ctx.sync(); ctx.prof_enable(True); ctx.prof_read(reset=True)
    t0 = time.perf_counter()
    K = 20
    for _ in range(K):
        ldb.query_global_device(dq, 2000, top)
    ctx.sync()
    dt = (time.perf_counter() - t0) / K * 1e3
    pr = {k: round(v[0] / K, 4) for k, v in ctx.prof_read(reset=True).items() if v[1]}
    ctx.prof_enable(False)
    out[rows] = {"ms": round(dt, 4), **pr}
    # the fused scan + exchange kernel with this GPU as its only peer (what every rank of an N-GPU run launches)
    from aria_slam_b200._lib import lib
    words = lib().aslam_loop_exchange_buffer_bytes(1, 2000) // 8
    buf = torch.zeros(words, dtype=torch.int64, device="cuda")
    ptrs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    rnd = 0
    for _ in range(3):
        rnd += 1; ldb.query_global_exchange_device(dq, 2000, ptrs, 1, 0, rnd, top)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.ExternalStream(ctx.stream)
    e0.record(st)
    for _ in range(K):
        rnd += 1; ldb.query_global_exchange_device(dq, 2000, ptrs, 1, 0, rnd, top)
    e1.record(st)
    ctx.sync()
    out[rows]["fused_ms"] = round(e0.elapsed_time(e1) / K, 4)
    ldb.close()
print(json.dumps(out))
