#!/usr/bin/env python
"""bench.py -- ORB extract+match frames/s @752x480 (BASELINE.json metric, config C2) on N B200s of one
node, plus the sharded loop-closure query (config C5) as a second figure on the same JSON line.

  python bench.py --gpus 1 --steps 10 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the reference's CPU path (cv2), same metric/config

A step = one pass of the hot path over one batch of 256 synthetic 752x480 frames per GPU: ORB
extraction (2000 keypoints) of every frame + Hamming kNN2/ratio matching of every frame against
its predecessor.  `value` = frames/s with the batch already resident in HBM; `e2e` = the same
through the C-ABI call with pinned host buffers, H2D of the frames and D2H of keypoints,
descriptors and matches inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NFEAT, BATCH = 752, 480, 2000, 256
METRIC = "orb_extract_match_frames_per_s_752x480"
WORKLOAD = "C2: stream of 256 synthetic 752x480 frames per GPU, ORB 2000 kpts (8 levels) + frame-to-frame Hamming kNN2 ratio 0.75"
# algorithmic bytes per 752x480 frame, each stage streaming its input once
BYTES_PER_FRAME = {"resize": 1_845_634, "fast_nms": 1_117_367, "blur": 2_234_734}
LOOP_DB_ROWS, LOOP_NQ, LOOP_KF_ROWS = 1_000_000, 2000, 2000


def bench_config():
    """The `config` object of both arms (ours and --impl reference): the same workload, the same 256 frames per GPU and step."""
    return {"workload": WORKLOAD, "frames_per_gpu_per_step": BATCH,
            "l2_policy": "working set per step (pyramid+blur+candidates ~1.4 GB) exceeds the 126 MB L2",
            "multi_gpu": "replicas: independent frame streams per GPU, no collective"}


def host_info():
    info = {"cores": os.cpu_count() or 1, "cpu": cpu_model()}
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["numa_nodes"] = len(nodes)
    except OSError:
        info["numa_nodes"] = None
    return info


def env_int(name, default):
    return int(os.environ.get(name, default))


_SAMPLER_SRC = r"""
import sys, time
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
period = float(sys.argv[2])
reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
print("max", pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM), flush=True)
while True:
    try:
        print(repr(time.time()), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), int(reasons(h)), flush=True)
    except Exception:
        pass
    time.sleep(period)
"""


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every 2 ms while the timed regions run.  The sampler is a
    separate PROCESS (a thread in this interpreter competes with the launching thread for the GIL and, at N = 8, made
    rank 0 the straggler every other rank's exchange waited for); its samples carry wall-clock stamps and are matched
    to the timed regions afterwards.  Falls back to the recipe's nvidia-smi line when pynvml is not importable."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index, period_s=0.002):
        self.index, self.period, self.rows, self.max_mhz = index, period_s, [], None
        self.windows, self._t0 = [], None      # wall-clock windows of the timed regions
        self.proc, self.source = None, None

    @property
    def timed(self):
        return self._t0 is not None

    @timed.setter
    def timed(self, on):
        if on and self._t0 is None:
            self._t0 = time.time()
        elif not on and self._t0 is not None:
            self.windows.append((self._t0, time.time()))
            self._t0 = None

    def start(self):
        try:
            import pynvml  # noqa: F401  (only to know the subprocess can import it)
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(self.index), str(self.period)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvml (separate process)"
        except Exception:
            self._start_smi()
        if self.proc:
            threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:        # blocks in read(): the GIL is released while waiting
            c = line.split()
            try:
                if c[0] == "max":
                    self.max_mhz = float(c[1])
                elif self.source.startswith("nvml"):
                    self.rows.append((float(c[0]), float(c[1]), int(c[2])))
                else:
                    c = [x.strip() for x in line.split(",")]
                    self.max_mhz = float(c[1])
                    why = sum(bit for (_, bit), v in zip(self.REASONS, c[2:6]) if v.lower().startswith("active"))
                    self.rows.append((time.time(), float(c[0]), why))
            except (ValueError, IndexError):
                pass

    def _start_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
        except OSError:
            self.proc = None

    def stop(self):
        self.timed = False
        if self.proc:
            self.proc.terminate()
        rows = list(self.rows)
        timed = [r for r in rows if any(a <= r[0] <= b for a, b in self.windows)] or rows
        sm = [r[1] for r in timed]
        bits = 0
        for r in timed:
            bits |= r[2]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(name for name, bit in self.REASONS if bits & bit),
                "samples": len(sm), "samples_total": len(rows), "timed_windows": len(self.windows), "source": self.source}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def bind_to_gpu_cpus(index):
    """Best effort: run this rank (and first-touch its pinned buffers) on the CPUs NVML reports as local
    to its GPU, so that H2D/D2H traffic of 8 ranks does not cross sockets."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
    except Exception:
        pass


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ----------------------------------------------------------------------------------- reference CPU arm
def cv2_extract_match(frames, nfeat, threads):
    """The reference's CPU path: cv::ORB::detectAndCompute (Frame.cpp:47) per frame, then
    BFMatcher(NORM_HAMMING).knnMatch(k=2) + 0.75 ratio (main.cpp:159-175) against the previous frame."""
    import cv2
    cv2.setNumThreads(threads)
    orb = cv2.ORB_create(nfeat, 1.2, 8, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 20)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    prev, good = None, 0
    for f in frames:
        _, d = orb.detectAndCompute(f, None)
        if prev is not None and d is not None and len(prev) >= 2:
            for m in bf.knnMatch(prev, d, k=2):          # knnMatch(prev_frame, current_frame), main.cpp:158-161
                if len(m) >= 2 and m[0].distance < 0.75 * m[1].distance:
                    good += 1
        prev = d
    return good


class Cv2Streams:
    """The same CPU path run the way the GPU arm runs it: the batch is cut into independent frame streams, one host
    thread each with one OpenCV thread (cv::ORB parallelises only resize/FAST/blur internally and scales poorly with
    cv2.setNumThreads; independent streams use the cores about 4x better).  Like the GPU arm's streams, a stream keeps
    its last frame's descriptors between steps, so every frame of a step is extracted once and matched once."""

    def __init__(self, frames, nfeat, workers):
        import cv2
        from concurrent.futures import ThreadPoolExecutor
        cv2.setNumThreads(1)
        self.cv2 = cv2
        n = len(frames)
        workers = max(1, min(workers, n))
        cuts = [n * k // workers for k in range(workers + 1)]
        self.chunks = [frames[cuts[k]:cuts[k + 1]] for k in range(workers)]
        self.prev = [None] * workers
        self.orbs = [cv2.ORB_create(nfeat, 1.2, 8, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 20) for _ in range(workers)]
        self.bfs = [cv2.BFMatcher(cv2.NORM_HAMMING) for _ in range(workers)]
        self.pool = ThreadPoolExecutor(workers)
        self.workers, self.frames_per_step = workers, n

    def _run(self, k, limit):
        orb, bf, prev, good = self.orbs[k], self.bfs[k], self.prev[k], 0
        for f in self.chunks[k][:limit]:
            _, d = orb.detectAndCompute(f, None)
            if prev is not None and d is not None and len(prev) >= 2 and len(d) >= 2:
                for m in bf.knnMatch(prev, d, k=2):          # knnMatch(prev_frame, current_frame), main.cpp:158-161
                    if len(m) >= 2 and m[0].distance < 0.75 * m[1].distance:
                        good += 1
            prev = d
        self.prev[k] = prev
        return good

    def step(self, limit=None):
        """One pass over the batch (or over the first `limit` frames of every stream); returns frames processed."""
        list(self.pool.map(lambda k: self._run(k, limit), range(self.workers)))
        return sum(len(c[:limit]) for c in self.chunks)

    def close(self):
        self.pool.shutdown()


def cpu_baseline_frames(frames, budget_s=12.0):
    cores = os.cpu_count() or 1
    import cv2
    cv2_extract_match(frames[:2], NFEAT, cores)          # warm-up
    t0 = time.perf_counter()
    cv2_extract_match(frames[:9], NFEAT, cores)
    single = 8 / (time.perf_counter() - t0)
    st = Cv2Streams(frames, NFEAT, cores)
    st.step(1)                                           # warm-up of the pool path
    per, n, t0 = 4, 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        n += st.step(per)
    dt = time.perf_counter() - t0
    st.close()
    return {"value": n / dt, "unit": "frames/s", "cores": st.workers, "kind": "reference",
            "sample": "%d frames (%d independent streams x %d frames per pass, repeated for %.0f s), cv2 %s ORB_create(2000)+BFMatcher "
                      "knnMatch(k=2)+0.75 ratio, one OpenCV thread per stream, %s" % (n, st.workers, per, dt, cv2.__version__, cpu_model()),
            "single_stream_all_cv2_threads": {"value": single, "unit": "frames/s", "threads": cores}}


def cpu_baseline_loop(db, q, rows=100_000):
    import cv2
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    sub = np.ascontiguousarray(db[:rows])
    bf.knnMatch(q[:64], sub[:10000], k=2)
    t0 = time.perf_counter()
    bf.knnMatch(q, sub, k=2)
    dt = time.perf_counter() - t0
    return {"value": len(q) / (dt * len(db) / rows), "unit": "query-descriptors/s", "cores": cores, "kind": "reference",
            "sample": "cv2 BFMatcher knnMatch(k=2) of %d queries against %d of the %d rows, time scaled x%d" % (len(q), rows, len(db), len(db) // rows)}


def single_frame_latency(ctx, frames, reps=50):
    """C1-like use: one 752x480 frame through IFeatureExtractor::extract (host image in, host keypoints and
    descriptors out) and one IMatcher::match against the previous frame, 1000 keypoints; wall clock."""
    from aria_slam_b200.frontend import OrbB200Extractor, B200Matcher, Frame
    ex, m = OrbB200Extractor(1000, ctx), B200Matcher(ctx)
    fa, fb = Frame(), Frame()
    for _ in range(5):
        ex.extract(frames[0], W, H, fa); ex.extract(frames[1], W, H, fb); mm = []; m.match(fb, fa, mm, 0.75)
    t0 = time.perf_counter()
    for _ in range(reps):
        ex.extract(frames[1], W, H, fb)
    t1 = time.perf_counter()
    for _ in range(reps):
        mm = []; m.match(fb, fa, mm, 0.75)
    t2 = time.perf_counter()
    ex.close()
    return {"config": "C1: 752x480, ORB 1000 kpts, kNN2 + 0.75 ratio, one frame per call", "extract_ms": (t1 - t0) / reps * 1e3,
            "match_ms": (t2 - t1) / reps * 1e3, "keypoints": fb.numKeypoints(), "matches": len(mm)}


def run_reference(args):
    """The reference's own CPU path (cv2 = the OpenCV the reference links) on this box's host cores, on OUR arm's config:
    every step is the full 256-frame batch, cut into one frame stream per core."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from aria_slam_b200 import synth
    cores = os.cpu_count() or 1
    frames = synth.frame_stream(BATCH, W, H, 1000)
    st = Cv2Streams(frames, NFEAT, cores)
    for _ in range(max(args.warmup, 1)):
        st.step(1)                                          # one frame per stream: thread pool, caches, cv2 dispatch
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st.step()
    dt = time.perf_counter() - t0
    val = BATCH * args.steps / dt
    sample = "the full %d-frame batch per step (%d independent streams, one OpenCV thread each); cv2 %s ORB+BFMatcher, %s" % (
        BATCH, st.workers, __import__("cv2").__version__, cpu_model())
    st.close()
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": bench_config(), "host": host_info(),
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": st.workers, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


# ----------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from aria_slam_b200 import synth
    from aria_slam_b200._lib import Context, kernel_launches, KP_DTYPE, MATCH_DTYPE
    from aria_slam_b200.frontend import OrbBatchTracker, LoopDB, unpack_top2
    from aria_slam_b200.sharded import ShardedLoopQuery

    world, rank, local = env_int("WORLD_SIZE", 1), env_int("RANK", 0), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback")
    torch.cuda.set_device(local)
    bind_to_gpu_cpus(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = torch.cuda.Stream()
    ctx = Context(local, stream.cuda_stream)
    K, Wm = args.steps, max(args.warmup, 3)
    frames = synth.frame_stream(BATCH, W, H, 1000 + rank * BATCH)
    cap_bytes = None

    # The 256-frame batch of a step is cut into independent frame streams (one handle + CUDA stream
    # each): kernels of different streams overlap, which fills the latency-bound phases (selection,
    # description, the small pyramid levels) of one stream with the ALU/POPC-bound kernels of another.
    n_streams = env_int("ASLAM_STREAMS", 8)
    half = BATCH // n_streams
    h_in = torch.from_numpy(frames).pin_memory()
    with torch.cuda.stream(stream):
        d_frames = torch.from_numpy(frames).cuda()
    pipes, pipe_streams = [], []
    for s in range(n_streams):
        st = torch.cuda.Stream()
        c = Context(local, st.cuda_stream)
        t = OrbBatchTracker(W, H, NFEAT, half, c)
        cap = t.capacity
        outs = dict(kps=torch.empty((half, cap, 24), dtype=torch.uint8).pin_memory(),
                    desc=torch.empty((half, cap, 32), dtype=torch.uint8).pin_memory(),
                    cnt=torch.empty(half, dtype=torch.int32).pin_memory(),
                    m=torch.empty((half, cap, 12), dtype=torch.uint8).pin_memory(),
                    mc=torch.empty(half, dtype=torch.int32).pin_memory())
        pipes.append((c, t, outs, h_in[s * half:(s + 1) * half]))
        pipe_streams.append(st)
    h2d = BATCH * W * H
    # the library copies nfeatures rows per frame up front (tie tails beyond that only when a frame has them) + 3 counters
    d2h = BATCH * min(cap, NFEAT) * (24 + 32 + 12) + BATCH * 12

    # ---- value: batch resident in HBM, all streams running, timed on the device (fork/join events)
    def value_run(steps):
        for it in range(steps):
            for s, (c, t, o, src) in enumerate(pipes):
                t.track_async(d_frames[s * half:(s + 1) * half], 0.75, on_device=True); t.device_results()

    torch.cuda.synchronize()
    value_run(Wm)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:                           # one sampler per job: nvidia-smi polling is not free on the host
        clocks.start()
    launches0 = kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.timed = True
    e0.record(stream)
    for st in pipe_streams:
        st.wait_event(e0)
    value_run(K)
    for st in pipe_streams:
        ev = torch.cuda.Event(); ev.record(st); stream.wait_event(ev)
    e1.record(stream)
    barrier()
    clocks.timed = False
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = kernel_launches() - launches0
    value = world * BATCH * K / (ms * 1e-3)

    # ---- per-kernel durations for the roofline: the same K steps replayed on ONE stream (kernels do
    # not overlap there, so an event pair around a launch times that kernel alone)
    with torch.cuda.stream(stream):
        tr = OrbBatchTracker(W, H, NFEAT, BATCH, ctx)
        for _ in range(Wm):
            tr.track_async(d_frames, 0.75, on_device=True); tr.device_results()
        ctx.prof_enable(True)
        barrier(); ctx.prof_read(reset=True)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(K):
            tr.track_async(d_frames, 0.75, on_device=True); tr.device_results()
        s1.record(stream)
        barrier()
        serial_ms = max_over_ranks(s0.elapsed_time(s1))
        prof = ctx.prof_read(reset=True)
        ctx.prof_enable(False)

    # ---- e2e: pinned host frames in, pinned results out, through the C-ABI host-buffer calls
    # Software pipeline over the independent frame streams: a stream's results are collected (D2H +
    # stream sync) and its next batch is submitted at once, so the other streams' copies and kernels
    # keep the GPU and both PCIe directions busy meanwhile.  K steps = K batches per stream.
    def e2e_run(steps):
        for c, t, o, src in pipes:
            t.track_async(src, 0.75, on_device=False)
        for it in range(steps):
            for c, t, o, src in pipes:
                t.sync(o["kps"], o["desc"], o["cnt"], o["m"], o["mc"])
                if it + 1 < steps:
                    t.track_async(src, 0.75, on_device=False)

    e2e_run(Wm)
    barrier()
    # timed on the device: the frame streams start behind one event and a last event waits for all of them,
    # so the span covers every H2D copy, kernel and D2H copy of the K steps; the wall clock is kept beside it
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    for st in pipe_streams:
        st.wait_event(g0)
    clocks.timed = True
    t0 = time.perf_counter()
    e2e_run(K)
    for st in pipe_streams:
        ev = torch.cuda.Event(); ev.record(st); stream.wait_event(ev)
    g1.record(stream)
    torch.cuda.synchronize()
    clocks.timed = False
    e2e_wall_s = max_over_ranks(time.perf_counter() - t0)
    e2e_s = max_over_ranks(g0.elapsed_time(g1) * 1e-3)
    e2e_val = world * BATCH * K / e2e_s
    cnt_np = np.concatenate([p[2]["cnt"].numpy() for p in pipes])
    mc_np = np.concatenate([p[2]["mc"].numpy() for p in pipes])
    checks = {"frames_at_quota": int((cnt_np == NFEAT).sum()), "median_matches_shifted_pairs": float(np.median(mc_np[1::2]))}

    # ---- copy-only probe: the step's H2D and D2H bytes with no kernels at all, every rank at once.  It separates "PCIe /
    # host memory cannot carry more" from "the GPU is waiting for something else" when end-to-end scaling falls off.
    def copy_only(reps=8):
        d_in = torch.empty((BATCH, H, W), dtype=torch.uint8, device="cuda")
        d_out = torch.empty(d2h, dtype=torch.uint8, device="cuda")
        h_out = torch.empty(d2h, dtype=torch.uint8).pin_memory()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        res = {}
        for name, up, down in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
            for timed in (False, True):
                barrier()
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record(stream); s_in.wait_event(c0); s_out.wait_event(c0)
                for _ in range(reps if timed else 2):
                    if up:
                        with torch.cuda.stream(s_in):
                            d_in.copy_(h_in, non_blocking=True)
                    if down:
                        with torch.cuda.stream(s_out):
                            h_out.copy_(d_out, non_blocking=True)
                for st_ in (s_in, s_out):
                    ev = torch.cuda.Event(); ev.record(st_); stream.wait_event(ev)
                c1.record(stream)
                barrier()
            sec = max_over_ranks(c0.elapsed_time(c1) * 1e-3) / reps
            res[name] = {"ms_per_step_bytes": sec * 1e3, "GBps_per_gpu": ((h2d if up else 0) + (d2h if down else 0)) / sec / 1e9,
                         "GBps_all_gpus": world * ((h2d if up else 0) + (d2h if down else 0)) / sec / 1e9}
        res["frames_per_s_if_copies_were_the_only_cost"] = world * BATCH / (res["both"]["ms_per_step_bytes"] * 1e-3)
        return res

    copy_probe = copy_only()

    # ---- loop-closure query, database sharded over the ranks (strong scaling)
    db = synth.descriptor_db(LOOP_DB_ROWS, 5)
    q, src = synth.loop_queries(db, LOOP_NQ, 6)
    nqb = (LOOP_NQ + 511) // 512

    def timed_loop(fn, steps, reset=None):
        """K calls of fn() back to back on `stream`, CUDA events, max over ranks -> ms per call."""
        for _ in range(Wm):
            fn()
        barrier()
        if reset:
            reset()
        t0_, t1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0_.record(stream)
        for _ in range(steps):
            fn()
        t1_.record(stream)
        barrier()
        return max_over_ranks(t0_.elapsed_time(t1_)) / steps

    with torch.cuda.stream(stream):
        # consecutive queries of the throughput loop are pipelined (programmatic dependent launch: the next query's scan
        # starts while the previous one's exchange tail runs); ASLAM_BENCH_PIPELINED=0 times them strictly one after another
        pipelined = env_int("ASLAM_BENCH_PIPELINED", 1) == 1
        lq = ShardedLoopQuery(ctx, db, rank, world, keyframe_rows=LOOP_KF_ROWS, pipelined=pipelined)
        d_q = torch.from_numpy(q).cuda()
        h_q = torch.from_numpy(q).pin_memory()
        h_top2 = torch.empty((LOOP_NQ, 2), dtype=torch.int64).pin_memory()
        lq.query(d_q, LOOP_NQ)
        barrier()
        # flat query, device-resident
        launches1 = kernel_launches()
        clocks.timed = True
        loop_ms = timed_loop(lambda: lq.query(d_q, LOOP_NQ), K, reset=lambda: lq.db.exchange_status(0, reset=True) if lq.exchange in ("ipc", "fused") else None)
        clocks.timed = False
        loop_launches = (kernel_launches() - launches1) * K // (K + Wm)
        # per-kernel durations from a second pass with an event pair around every launch (the pairs keep consecutive
        # queries apart, so this pass is not pipelined and is not the one the throughput is taken from)
        wait_us = lq.db.mean_exchange_wait_us(nqb) if lq.exchange in ("ipc", "fused") else None     # mean over the K timed queries
        ctx.prof_enable(True); ctx.prof_read(reset=True)
        loop_serial_ms = timed_loop(lambda: lq.query(d_q, LOOP_NQ), K)
        loop_prof = {k: (v[0] * K / (K + Wm), v[1] * K // (K + Wm)) for k, v in ctx.prof_read(reset=True).items()}
        ctx.prof_enable(False)
        top2 = lq.query(d_q, LOOP_NQ)
        # flat query end to end: pinned host queries in (64 KB), merged top-2 out to pinned host memory (32 KB)
        d_q2 = torch.empty_like(d_q)

        def flat_e2e():
            d_q2.copy_(h_q, non_blocking=True)
            h_top2.copy_(lq.query(d_q2, LOOP_NQ), non_blocking=True)
        loop_e2e_ms = timed_loop(flat_e2e, K)
        # the reference-shaped query: ratio-0.7 survivors per keyframe (findCandidates, LoopClosure.cpp:86-95), sharded by
        # keyframe; the exchange is one all-gather of int32 counts
        seg_ms = timed_loop(lambda: lq.query_segmented(d_q, LOOP_NQ), K)
        good = lq.query_segmented(d_q, LOOP_NQ)
        h_good = torch.empty(lq.n_keyframes, dtype=torch.int32).pin_memory()

        def seg_e2e():
            d_q2.copy_(h_q, non_blocking=True)
            h_good.copy_(lq.query_segmented(d_q2, LOOP_NQ), non_blocking=True)
        seg_e2e_ms = timed_loop(seg_e2e, K)
        torch.cuda.synchronize()
    clk = clocks.stop()
    keys = top2.cpu().numpy().view(np.uint64).copy()
    good_np = good.cpu().numpy().copy()

    # ---- parity of the sharded queries, checked inside the bench at every N (the full [nq, 2] keys: both neighbours and
    # both distances, which is what a faulty merge or flag protocol would corrupt)
    loop_checks = {}
    if world > 1:
        gathered = torch.empty((world, LOOP_NQ, 2), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(gathered.view(-1, 2), top2.contiguous())
        loop_checks["ranks_identical"] = bool((gathered == gathered[0]).all().item())
        gg = torch.empty((world, lq.n_keyframes), dtype=torch.int32, device="cuda")
        dist.all_gather_into_tensor(gg.view(-1), good.contiguous())
        loop_checks["segmented_ranks_identical"] = bool((gg == gg[0]).all().item())
    else:
        loop_checks["ranks_identical"] = True
        loop_checks["segmented_ranks_identical"] = True
    if rank == 0:
        import cv2
        # the whole database on ONE GPU (rank 0), before any sharding
        whole = LoopDB(ctx, LOOP_DB_ROWS, LOOP_DB_ROWS // LOOP_KF_ROWS)
        for k in range(LOOP_DB_ROWS // LOOP_KF_ROWS):
            whole.add(db[k * LOOP_KF_ROWS:(k + 1) * LOOP_KF_ROWS], k)
        one = np.zeros((LOOP_NQ, 2), np.uint64)
        from aria_slam_b200._lib import check, lib, ptr
        check(lib().aslam_loopdb_query_global(whole.h, ptr(q), LOOP_NQ, ptr(one)))
        loop_checks["equals_single_shard"] = bool((keys == one).all())
        loop_checks["segmented_equals_single_gpu"] = bool((good_np == whole.query_segmented(q, 0.7)).all())
        whole.close()
        # a sample of queries against the reference's own matcher over all 1 M rows
        cv2.setNumThreads(os.cpu_count() or 1)
        pick = np.arange(0, LOOP_NQ, LOOP_NQ // 80)[:80]
        bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        # cv2 refuses train sets of 2^18 rows or more (IMGIDX_ONE): match in blocks of 250 000 rows, keep the two smallest
        # (distance, row) per query -- the order cv2 uses inside a block
        best = [[] for _ in pick]
        for lo in range(0, LOOP_DB_ROWS, 250_000):
            for i, m in enumerate(bf.knnMatch(q[pick], db[lo:lo + 250_000], k=2)):
                best[i] += [(int(e.distance), lo + e.trainIdx) for e in m]
        ok = True
        for cand, row in zip(best, keys[pick]):
            r_, d_ = unpack_top2(row)
            two = sorted(cand)[:2]
            ok = ok and len(two) == 2 and [two[0][1], two[1][1]] == list(r_) and [two[0][0], two[1][0]] == list(d_)
        loop_checks["sample_equals_cv2"] = bool(ok)
        loop_checks["sample_queries"] = int(len(pick))
        loop_checks["second_neighbour_checked"] = True
        # per-keyframe counts against cv2 on a sample of keyframes (the query's source keyframes are among them)
        kf_pick = sorted(set([0, 1, 250, 499] + [int(r) // LOOP_KF_ROWS for r in src[:4]]))
        okc = True
        for kf in kf_pick:
            cnt = 0
            for m in bf.knnMatch(q, db[kf * LOOP_KF_ROWS:(kf + 1) * LOOP_KF_ROWS], k=2):
                if len(m) >= 2 and m[0].distance < 0.7 * m[1].distance:       # LoopClosure.cpp:92
                    cnt += 1
            okc = okc and cnt == int(good_np[kf])
        loop_checks["segmented_sample_equals_cv2"] = bool(okc)
        loop_checks["segmented_sample_keyframes"] = len(kf_pick)
        loop_checks["keys_crc32"] = int(__import__("zlib").crc32(keys.tobytes()))
        loop_checks["exchange_status_ok"] = wait_us is None or wait_us >= 0
    rows, dist_ = unpack_top2(keys)
    loop_ok = bool((rows[:, 0] == src).all())
    loop_val = LOOP_NQ * K / (loop_ms * K * 1e-3)
    waits = None
    if world > 1 and wait_us is not None:
        wt = torch.tensor([wait_us], dtype=torch.float64, device="cuda")
        allw = torch.empty(world, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allw, wt)
        waits = [round(x, 1) for x in allw.cpu().tolist()]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel + all stages
    hbm_peak, peak_src = measured_peaks()
    popc_peak = ctx.microbench(0)                      # G popc32/s, measured on this GPU now
    frames_total = BATCH * K
    stages = {}
    step_ms = sum(v[0] for v in prof.values()) / K
    pairs_per_step = float(np.sum(cnt_np[1:].astype(np.float64) * cnt_np[:-1])) + 0.0
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    pipes = {}
    pp = os.path.join(ROOT, "profiles", "pipes.json")
    if os.path.exists(pp):
        pipes = {k: dict(v, source=json.load(open(pp)).get("_source")) for k, v in json.load(open(pp)).items() if not k.startswith("_")}
    POPC_EXECUTED, POPC_ALGORITHMIC = 4.0, 8.0         # per descriptor pair: the carry-save form executes 4 (match.cu), SURVEY 8d counts 8
    for name, (tot_ms, n) in prof.items():
        if n == 0:
            continue
        ent = {"ms_per_step": tot_ms / K, "launches_per_step": n / K, "share": (tot_ms / K) / step_ms}
        if name in BYTES_PER_FRAME:
            ach = BYTES_PER_FRAME[name] * frames_total / (tot_ms * 1e-3) / 1e9
            ent.update(bound="hbm", achieved=ach, peak=hbm_peak, unit="GB/s", frac=ach / hbm_peak)
        elif name == "knn2_partial":
            # `frac` = POPC instructions the kernel EXECUTES (4 per pair) over the POPC rate measured on this GPU;
            # the algorithmic count of SURVEY 8d (8 per pair) is reported beside it as frac_algorithmic
            ach = POPC_EXECUTED * pairs_per_step * K / (tot_ms * 1e-3) / 1e9
            ent.update(bound="int_pipe(popc)", achieved=ach, peak=popc_peak, unit="Gpopc32/s executed", frac=ach / popc_peak,
                       executed_popc_per_pair=POPC_EXECUTED, algorithmic_popc_per_pair=POPC_ALGORITHMIC,
                       frac_algorithmic=POPC_ALGORITHMIC / POPC_EXECUTED * ach / popc_peak)
            # the same launch seen as HBM traffic (each descriptor row read once: 32 B per query and per train row) --
            # far below the HBM roofline, which is why the bound above is the POPC pipe
            kb = 32.0 * float(np.sum(cnt_np[1:].astype(np.float64) + cnt_np[:-1])) * K / (tot_ms * 1e-3) / 1e9
            ent["hbm_view"] = {"bound": "hbm", "achieved": kb, "peak": hbm_peak, "unit": "GB/s", "frac": kb / hbm_peak}
        ent["traffic"] = traffic.get(name)
        if name in pipes:
            ent["ncu"] = pipes[name]        # pipe / memory utilisation of this kernel kind from the committed ncu capture
        stages[name] = ent
    dom = max(stages, key=lambda k: stages[k]["ms_per_step"])
    roof = dict(stages[dom]); roof["kernel"] = dom; roof["peak_source"] = peak_src if roof.get("bound") == "hbm" else "aslam_microbench(POPC) on this GPU (MEASURED_PEAKS.json has no integer-pipe entry)"
    for k in ("bound", "achieved", "peak", "unit", "frac"):
        roof.setdefault(k, None)
    grp_ms = sum(stages[k]["ms_per_step"] for k in ("resize", "fast_nms") if k in stages)
    grp = (BYTES_PER_FRAME["resize"] + BYTES_PER_FRAME["fast_nms"]) * BATCH / (grp_ms * 1e-3) / 1e9 if grp_ms else None

    # loop query: its own roofline (shard scan = XOR/POPC over nq x rows/N pairs per GPU) and end-to-end figure
    scan_ms = loop_prof.get("knn2_partial", (0.0, 0))[0] / K
    pairs_per_gpu = LOOP_NQ * float(LOOP_DB_ROWS) / world
    loop_roof = None
    if scan_ms > 0:
        ach = POPC_EXECUTED * pairs_per_gpu / (scan_ms * 1e-3) / 1e9
        loop_roof = {"kernel": "knn2_partial (shard scan%s)" % (" + fused exchange tail" if lq.exchange in ("ipc", "fused") else ""),
                     "bound": "int_pipe(popc)", "achieved": ach, "peak": popc_peak, "unit": "Gpopc32/s executed per GPU", "frac": ach / popc_peak,
                     "frac_algorithmic": POPC_ALGORITHMIC / POPC_EXECUTED * ach / popc_peak, "ms": scan_ms,
                     "peak_source": "aslam_microbench(POPC) on this GPU",
                     "note": "at N > 1 the fused kernel's duration includes the wait for the slowest peer (mean_wait_for_slowest_peer_us_per_rank)"}

    out = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": bench_config(), "streams_per_gpu": n_streams, "host": host_info(),
        "clocks": clk, "gpu_launches": int(launches),
        "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "streams_per_gpu": n_streams, "timing": "CUDA events around all frame streams, max over ranks",
                "wall_clock_value": world * BATCH * K / e2e_wall_s,
                "copy_only_probe": copy_probe},
        "single_stream": {"ms_per_step": serial_ms / K, "value": world * BATCH * K / (serial_ms * 1e-3),
                          "note": "the K steps replayed on one stream with an event pair around every launch; roofline durations come from here"},
        "roofline": roof, "roofline_stages": stages,
        "roofline_group_pyramid_fast": {"bound": "hbm", "achieved": grp, "peak": hbm_peak, "unit": "GB/s", "frac": grp / hbm_peak if grp else None},
        "checks": checks,
        "loop_query": {"metric": "loop_query_descriptors_per_s", "value": loop_val, "unit": "query-descriptors/s", "db_rows": LOOP_DB_ROWS,
                       "queries": LOOP_NQ, "keyframes": lq.n_keyframes, "n_gpus": world, "scaling": "strong", "ms_per_query_batch": loop_ms,
                       "collective": {"none": "none (1 GPU)",
                                      "ipc": "none: ONE kernel per query -- the shard scan whose last CTA per 512-query block stores the block's packed (dist,row) keys into every peer's buffer over NVLink (cudaMalloc buffers opened through CUDA IPC by the library), raises release flags, acquires the peers' flags and merges",
                                      "fused": "as ipc, over buffers from torch symmetric memory",
                                      "push": "P2P stores from the finish kernel + symmetric-memory barriers + merge kernel",
                                      "nccl": "NCCL all_gather of 2000x2 packed (dist,row) keys per rank + merge kernel"}[lq.exchange],
                       "exchange": lq.exchange, "mean_wait_for_slowest_peer_us_per_rank": waits,
                       "pipelined_consecutive_queries": bool(lq.pipelined and lq.exchange in ("ipc", "fused")),
                       "ms_per_query_batch_not_pipelined": loop_serial_ms,
                       "best_row_is_source_row": loop_ok, "checks": loop_checks, "gpu_launches_per_query": loop_launches / K,
                       "stages_ms": {k: v[0] / K for k, v in loop_prof.items() if v[1]},
                       "pairs_per_s": LOOP_NQ * float(LOOP_DB_ROWS) / (loop_ms * 1e-3),
                       "roofline": loop_roof,
                       "e2e": {"value": LOOP_NQ / (loop_e2e_ms * 1e-3), "unit": "query-descriptors/s", "ms_per_query_batch": loop_e2e_ms,
                               "h2d_bytes_per_step": LOOP_NQ * 32, "d2h_bytes_per_step": LOOP_NQ * 16,
                               "note": "pinned host queries in, merged top-2 keys out to pinned host memory, every query"},
                       "per_keyframe": {"what": "findCandidates' ratio-0.7 survivor count per stored keyframe (LoopClosure.cpp:86-95), database sharded by keyframe, counts all-gathered (4 B per keyframe)",
                                        "value": LOOP_NQ / (seg_ms * 1e-3), "unit": "query-descriptors/s", "ms_per_query_batch": seg_ms,
                                        "e2e": {"value": LOOP_NQ / (seg_e2e_ms * 1e-3), "unit": "query-descriptors/s", "ms_per_query_batch": seg_e2e_ms,
                                                "h2d_bytes_per_step": LOOP_NQ * 32, "d2h_bytes_per_step": lq.n_keyframes * 4}}},
    }
    if world == 1:
        out["latency_single_frame"] = single_frame_latency(ctx, frames)
        out["cpu_baseline"] = cpu_baseline_frames(frames)
        out["loop_query"]["cpu_baseline"] = cpu_baseline_loop(db, q)
    emit(out)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def claim_stdout():
    """The driver reads ONE JSON line from stdout.  Libraries (NCCL prints its version banner there)
    must not share it: fd 1 is pointed at stderr for the whole run, the result goes to the saved fd."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(obj):
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
