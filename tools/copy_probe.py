"""Run under torchrun (N ranks, one per GPU): host <-> device copy bandwidth of one bench step's bytes with every rank
copying at once -- plain pinned memory against write-combined pinned memory for the upload, downloads alone, and both
directions together.  Separates "PCIe / host memory cannot carry more" from anything the GPUs do (no kernel runs)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from aria_slam_b200._lib import Context, lib, check

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H2D, D2H, REPS = 256 * 752 * 480, 256 * 2000 * 68, 10
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
c_in, c_out = Context(local, s_in.cuda_stream), Context(local, s_out.cuda_stream)
d_in = torch.empty(H2D, dtype=torch.uint8, device="cuda"); d_out = torch.empty(D2H, dtype=torch.uint8, device="cuda")


def host(nbytes, wc):
    p = C.c_void_p()
    check((lib().aslam_host_alloc_wc if wc else lib().aslam_host_alloc)(nbytes, C.byref(p)))
    a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))
    a[:] = 7                                  # first touch (and the only kind of access write-combined memory likes)
    return p


h_pin, h_wc, h_out = host(H2D, False), host(H2D, True), host(D2H, False)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(up, down):
    res = 0.0
    for timed in (False, True):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        e0.record(cur); s_in.wait_event(e0); s_out.wait_event(e0)
        for _ in range(REPS if timed else 2):
            if up is not None:
                check(lib().aslam_memcpy_async(c_in.h, C.c_void_p(d_in.data_ptr()), up, H2D))
            if down:
                check(lib().aslam_memcpy_async(c_out.h, h_out, C.c_void_p(d_out.data_ptr()), D2H))
        for st in (s_in, s_out):
            ev = torch.cuda.Event(); ev.record(st); cur.wait_event(ev)
        e1.record(cur)
        barrier()
        res = e0.elapsed_time(e1) / REPS
    t = torch.tensor([res], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    nbytes = (H2D if up is not None else 0) + (D2H if down else 0)
    return {"ms": round(ms, 3), "GBps_per_gpu": round(nbytes / ms / 1e6, 1), "GBps_all_gpus": round(world * nbytes / ms / 1e6, 1)}


out = {"world": world, "host_cores": os.cpu_count(), "h2d_bytes": H2D, "d2h_bytes": D2H,
       "h2d_pinned": run(h_pin, False), "h2d_write_combined": run(h_wc, False), "d2h_pinned": run(None, True),
       "both_pinned": run(h_pin, True), "both_write_combined_upload": run(h_wc, True)}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
