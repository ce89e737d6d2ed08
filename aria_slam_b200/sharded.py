"""Loop-closure database sharded over the GPUs of one node.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  The keyframe database is cut
by keyframe into contiguous row ranges, one per rank; the 64 KB query block is replicated; every
rank scans its shard with the Hamming kernel and reports the two best (distance, GLOBAL row) per
query as packed 64-bit keys; one all-gather of 2000 x 2 x 8 B = 32 KB per rank follows and a
merge kernel takes the two smallest keys per query.  The key order is total, so the result is
bit-identical for every shard count (and equal to cv::BFMatcher on the whole database).

Frame streams (extract + frame-to-frame match) need no communication at all: they run as
independent replicas, one handle per GPU.
"""
import numpy as np


def plan_shards(n_keyframes, world):
    """Contiguous keyframe ranges [(lo, hi)) per rank, sizes differing by at most one."""
    base, extra = divmod(n_keyframes, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def keyframe_starts(total_rows, keyframe_rows):
    """Row offsets of keyframes of `keyframe_rows` rows each (the last may be shorter); n+1 entries."""
    s = list(range(0, total_rows, keyframe_rows)) + [total_rows]
    return np.asarray(s, np.int64)


def pack_top2(rows, dist):
    """(row[nq,2], dist[nq,2]) with -1 for empty -> packed keys dist << 40 | row (uint64)."""
    rows = np.asarray(rows, np.int64); dist = np.asarray(dist, np.int64)
    key = (dist.astype(np.uint64) << np.uint64(40)) | rows.astype(np.uint64)
    key[rows < 0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    return key


def all_gather_top2(local, world):
    """local: int64 tensor [nq, 2] of packed keys (device for NCCL, host for gloo) -> [world, nq, 2]."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local.unsqueeze(0)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())          # rank-major concatenation along dim 0
    return out.view((world,) + tuple(local.shape))


class ShardedLoopQuery:
    """This rank's shard of a flat descriptor database + the collective query.

    exchange = "fused": ONE kernel after the shard scan finishes it, stores the keys into every peer's
    buffer over NVLink (buffers from torch's symmetric memory), raises a release flag, acquires the
    peers' flags and merges -- no collective call and no separate barrier.
    exchange = "push": the same stores, but symmetric-memory barriers and a separate merge kernel.
    exchange = "nccl": NCCL all_gather_into_tensor + merge kernel.  "auto" tries fused, then nccl.
    """

    def __init__(self, ctx, db_rows, rank, world, keyframe_rows=2000, exchange="auto"):
        import torch
        from .frontend import LoopDB
        self.ctx, self.rank, self.world = ctx, rank, world
        starts = keyframe_starts(len(db_rows), keyframe_rows)
        lo_kf, hi_kf = plan_shards(len(starts) - 1, world)[rank]
        self.row_lo, self.row_hi = int(starts[lo_kf]), int(starts[hi_kf])
        self.db = LoopDB(ctx, max(self.row_hi - self.row_lo, 1), max(hi_kf - lo_kf, 1))
        for k in range(lo_kf, hi_kf):
            self.db.add(db_rows[starts[k]:starts[k + 1]], k)
        self.db.set_row_base(self.row_lo)
        self._torch = torch
        self._local = self._out = None
        self._symm = None
        self.exchange = "none" if world == 1 else exchange
        self._nq = 0

    def _setup_push(self, nq, device):
        """Symmetric [world, nq, 2] gather buffer on every rank + the table of peer pointers."""
        torch = self._torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from ._lib import lib
        words = lib().aslam_loop_exchange_buffer_bytes(self.world, nq) // 8     # data[2][world][nq][2] + flags[2][world]
        buf = symm_mem.empty((words,), dtype=torch.int64, device=device)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, dist.group.WORLD.group_name)
        ptrs = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64, device=device)
        torch.cuda.synchronize()
        dist.barrier()                      # every rank's flags are zero before anyone starts round 1
        self._symm = (buf, hdl, ptrs)
        self._round = 0

    def query(self, d_q, nq):
        """d_q: device tensor [nq, 32] u8 (replicated).  Returns a device int64 tensor [nq, 2] of packed
        keys, identical on every rank.  Runs on the current torch stream == the context's stream."""
        torch = self._torch
        from ._lib import lib, check, ptr
        if self._local is None or self._nq != nq:
            self._local = torch.empty((nq, 2), dtype=torch.int64, device=d_q.device)
            self._out = torch.empty((nq, 2), dtype=torch.int64, device=d_q.device)
            self._nq = nq
            if self.exchange in ("auto", "push", "fused"):
                try:
                    self._setup_push(nq, d_q.device)
                    if self.exchange == "auto":
                        self.exchange = "fused"
                except Exception:
                    if self.exchange != "auto":
                        raise
                    self.exchange = "nccl"
        if self.world == 1:
            self.db.query_global_device(d_q, nq, self._local)
            return self._local
        if self.exchange == "fused":
            buf, hdl, ptrs = self._symm
            self._round += 1
            self.db.query_global_exchange_device(d_q, nq, ptrs, self.world, self.rank, self._round, self._out)
            return self._out
        if self.exchange == "push":
            buf, hdl, ptrs = self._symm
            hdl.barrier(channel=0)          # every rank has consumed the previous round's buffer
            self.db.query_global_push_device(d_q, nq, ptrs, self.world, self.rank)
            hdl.barrier(channel=1)          # every shard has landed in every buffer
            parts = buf[:self.world * nq * 2].view(self.world, nq, 2)
        else:
            self.db.query_global_device(d_q, nq, self._local)
            parts = all_gather_top2(self._local, self.world)
        check(lib().aslam_loop_merge_top2_device(self.ctx.h, ptr(parts), self.world, nq, ptr(self._out)))
        return self._out
