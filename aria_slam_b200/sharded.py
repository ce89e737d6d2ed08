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


def all_gather_counts(local, kf_ranges, scratch=None):
    """Per-keyframe counts of every rank's shard -> one vector over all keyframes in global order.
    local: int32 tensor padded to the largest shard's keyframe count (device for NCCL, host for gloo);
    kf_ranges: plan_shards(...).  The only exchange of the sharded per-keyframe query."""
    import torch
    import torch.distributed as dist
    world = len(kf_ranges)
    per = local.shape[0]
    if world == 1:
        return local[:kf_ranges[0][1] - kf_ranges[0][0]]
    allc = scratch if scratch is not None else torch.empty(per * world, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(allc, local.contiguous())
    if all(hi - lo == per for lo, hi in kf_ranges):
        return allc
    return torch.cat([allc[r * per:r * per + (hi - lo)] for r, (lo, hi) in enumerate(kf_ranges)])


def _all_agree(ok, world):
    """True only if `ok` holds on EVERY rank (one all_reduce(MIN)): ranks must never commit to different exchange modes --
    a rank calling all_gather while its peers spin in the fused kernel would hang the job."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return bool(ok)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


class ShardedLoopQuery:
    """This rank's shard of a flat descriptor database + the collective queries (one process per GPU).

    Flat query (global top-2 per query descriptor):
      exchange = "ipc":   ONE kernel launch per query: the shard scan whose last CTA per query block stores the keys into
                          every peer's buffer over NVLink, raises a release flag, acquires the peers' flags and merges.
                          Peer buffers are cudaMalloc memory opened through CUDA IPC by the library itself
                          (aslam_loopdb_shard_export / aslam_loopdb_shard_init); only the 64-byte handles travel through
                          torch.distributed.
      exchange = "fused": the same kernel over buffers from torch's symmetric memory.
      exchange = "push":  P2P stores from a finish kernel + symmetric-memory barriers + merge kernel.
      exchange = "nccl":  NCCL all_gather_into_tensor + merge kernel.
      "auto" tries ipc, then fused, then nccl; the choice is agreed on by all ranks (all_reduce MIN) before anyone uses it.
    Per-keyframe query (the reference's findCandidates counts, LoopClosure.cpp:86-95): every rank counts ratio-test
    survivors for its own keyframes; the exchange is one all_gather of int32 counts.
    """

    def __init__(self, ctx, db_rows, rank, world, keyframe_rows=2000, exchange="auto", pipelined=False):
        import torch
        from .frontend import LoopDB
        self.ctx, self.rank, self.world = ctx, rank, world
        starts = keyframe_starts(len(db_rows), keyframe_rows)
        self.n_keyframes = len(starts) - 1
        self.kf_ranges = plan_shards(self.n_keyframes, world)
        lo_kf, hi_kf = self.kf_ranges[rank]
        self.row_lo, self.row_hi = int(starts[lo_kf]), int(starts[hi_kf])
        self.db = LoopDB(ctx, max(self.row_hi - self.row_lo, 1), max(hi_kf - lo_kf, 1))
        for k in range(lo_kf, hi_kf):
            self.db.add(db_rows[starts[k]:starts[k + 1]], k)
        self.db.set_row_base(self.row_lo)
        # consecutive fused queries may overlap (the next scan starts during the previous exchange tail): the caller
        # promises that no kernel writes the query buffer between two calls (aslam_loopdb_set_pipelining)
        self.pipelined = bool(pipelined)
        if self.pipelined:
            self.db.set_pipelining(True)
        self._torch = torch
        self._local = self._out = None
        self._symm = None
        self.exchange = "none" if world == 1 else exchange
        self._nq = 0
        self._good_local = self._good_all = None

    def _setup_symm(self, nq, device):
        """Symmetric exchange buffer on every rank + the table of peer pointers."""
        torch = self._torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from ._lib import lib
        words = lib().aslam_loop_exchange_buffer_bytes(self.world, nq) // 8     # data[2][world][nq][2] + flags[2][world][qblocks]
        buf = symm_mem.empty((words,), dtype=torch.int64, device=device)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, dist.group.WORLD.group_name)
        ptrs = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64, device=device)
        torch.cuda.synchronize()
        dist.barrier()                      # every rank's flags are zero before anyone starts round 1
        self._symm = (buf, hdl, ptrs)
        self._round = 0

    def _setup_ipc(self, nq):
        """Exchange buffers through the C ABI: export a CUDA IPC handle, all-gather the handles, open the peers'."""
        torch = self._torch
        import torch.distributed as dist
        handle = self.db.shard_export(self.world, nq)
        mine = torch.from_numpy(handle.copy()).cuda() if dist.get_backend() == "nccl" else torch.from_numpy(handle.copy())
        allh = torch.empty((self.world * len(handle),), dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(allh, mine)
        self.db.shard_init(self.rank, self.world, allh.cpu().numpy())

    def _choose_exchange(self, nq, device):
        want = self.exchange
        if want in ("auto", "ipc"):
            try:
                self._setup_ipc(nq)
                ok = True
            except Exception:
                if want == "ipc":
                    raise
                ok = False
            if _all_agree(ok, self.world):
                self._torch.cuda.synchronize()
                import torch.distributed as dist
                dist.barrier()              # every rank has opened every peer's (zeroed) buffer
                self.exchange = "ipc"
                return
            if want == "ipc":
                raise RuntimeError("CUDA IPC exchange buffers could not be opened on every rank")
        if want in ("auto", "fused", "push"):
            try:
                self._setup_symm(nq, device)
                ok = True
            except Exception:
                if want != "auto":
                    raise
                ok = False
            if _all_agree(ok, self.world):
                self.exchange = "fused" if want == "auto" else want
                return
            if want != "auto":
                raise RuntimeError("symmetric-memory exchange buffers could not be set up on every rank")
        self.exchange = "nccl"

    def query(self, d_q, nq):
        """d_q: device tensor [nq, 32] u8 (replicated).  Returns a device int64 tensor [nq, 2] of packed
        keys, identical on every rank.  Runs on the current torch stream == the context's stream."""
        torch = self._torch
        from ._lib import lib, check, ptr
        if self._local is None or self._nq != nq:
            if self._nq and self.exchange in ("ipc", "fused", "push"):
                raise ValueError("the exchange buffers were laid out for %d queries; use a new ShardedLoopQuery for %d" % (self._nq, nq))
            self._local = torch.empty((nq, 2), dtype=torch.int64, device=d_q.device)
            self._out = torch.empty((nq, 2), dtype=torch.int64, device=d_q.device)
            self._nq = nq
            if self.world > 1:
                self._choose_exchange(nq, d_q.device)
        if self.world == 1:
            self.db.query_global_device(d_q, nq, self._local)
            return self._local
        if self.exchange == "ipc":
            self.db.query_global_sharded_device(d_q, nq, self._out)
            return self._out
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

    def query_segmented(self, d_q, nq, ratio=0.7):
        """findCandidates' per-keyframe survivor counts over the whole sharded database: a device int32 tensor
        [n_keyframes] (keyframes in global order), identical on every rank.  Each rank counts for its own keyframes
        (aslam_loopdb_query_segmented_device); the exchange is an all_gather of the counts alone (4 B per keyframe)."""
        torch = self._torch
        per = max(max(hi - lo for lo, hi in self.kf_ranges), 1)
        if self._good_local is None:
            self._good_local = torch.zeros(per, dtype=torch.int32, device=d_q.device)
            self._good_all = torch.zeros(per * self.world, dtype=torch.int32, device=d_q.device)
        self.db.query_segmented_device(d_q, nq, self._good_local, ratio)
        return all_gather_counts(self._good_local, self.kf_ranges, self._good_all)
