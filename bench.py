#!/usr/bin/env python
"""bench.py -- Mevents/s through the event front-end (SAE + time surface + Arc* + temporal and
stereo LK) on synthetic stereo event streams (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (libesvio_fe.so)
  python bench.py --impl reference --gpus N --steps K ...  the CPU path on the host cores

A "step" is one window (1/30 s of stream time) of one stereo pair: both cameras' events go
through createSAE/time surface, the left camera through corner detection, then temporal and
stereo LK (FeatureTracker::trackEvent, feature_tracker/src/feature_tracker.cpp:340-603).
N > 1: every rank runs its own independent stereo stream (weak scaling, SURVEY.md 8e
"independent streams") and the ranks all-gather their packed track records once per window
over NCCL.  One JSON line is printed by rank 0.

At N = 1 the line also carries `batched` (a group of streams on the one GPU) and `frames`
(FeatureTracker::trackImage on synthetic stereo frames).

  torchrun --nproc-per-node 2 bench.py --split-lr [--workload W]   one stream split by camera
      over 2 GPUs, with the same windows on one GPU timed beside it
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from esvio_b200 import shard, synth  # noqa: E402

METRIC = "Mevents/s through time-surface+stereo LK"
UNIT = "Mevents/s"
DEFAULT_WORKLOAD = "stereo_davis346_1mevs"  # BASELINE.json configs[1]
L2_BYTES = 126 * 1024 * 1024


def env_int(k, d):
    return int(os.environ.get(k, d))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_sae_update_ts launch, from the
    committed `ncu --set full` capture of this workload (profiles/r1_k1_ncu_full.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_k1_ncu_full.json")) as f:
            return int(json.load(f)[workload]["traffic_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_cfg(name):
    w = synth.WORKLOADS[name]
    cfg = synth.default_config(w["width"], w["height"], max_cnt=w["max_cnt"],
                               min_dist=w["min_dist"], use_ransac=1)
    pub_div = int(round(synth.WINDOWS_PER_SEC / w["freq"]))  # freq 15 -> every 2nd window
    return w, cfg, pub_div


def gen_windows(w, stream, n):
    s = synth.StereoEventStream(w["width"], w["height"], w["rate"], stream=stream, mono=w["mono"])
    return [s.stereo_window(k) for k in range(n)]


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference C++ with real OpenCV (cv2) for LK/RANSAC
# ------------------------------------------------------------------------------------------
def cpu_tracker(cfg, threads):
    from oracle import oracle as ora  # the only place bench.py touches oracle/: the CPU legs
    ora.build()
    use_cv2 = ora.have_cv2()
    t = ora.OracleTracker(cfg, use_cv2=use_cv2, cv2_threads=threads)
    kind = "port"
    desc = ("oracle/ C restatement of event_detector.cc + feature_tracker.cpp, "
            + ("OpenCV %s calcOpticalFlowPyrLK/findFundamentalMat via cv2" % __import__("cv2").__version__
               if use_cv2 else "C port of OpenCV LK/RANSAC"))
    return t, kind, desc


def run_cpu(cfg, pub_div, wins, warmup, threads):
    t, kind, desc = cpu_tracker(cfg, threads)
    outs = []
    n_ev = 0
    t_total = 0.0
    for k, (L, R, tc) in enumerate(wins):
        t0 = time.perf_counter()
        o = t.track(tc, L, R, k % pub_div == 0)
        dt = time.perf_counter() - t0
        if k >= warmup:
            t_total += dt
            n_ev += len(L[0]) + len(R[0])
        outs.append(o)
    return n_ev, t_total, outs, kind, desc, t.timers()


def main_reference(args):
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return
    w, cfg, pub_div = workload_cfg(args.workload)
    cores = os.cpu_count() or 1
    wins = gen_windows(w, 0, args.steps + args.warmup)
    n_ev, sec, _, kind, desc, timers = run_cpu(cfg, pub_div, wins, args.warmup, cores)
    val = n_ev / sec / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "width": w["width"], "height": w["height"],
                   "events_per_window_per_camera": int(round(w["rate"] / synth.WINDOWS_PER_SEC)),
                   "max_cnt": w["max_cnt"], "min_dist": w["min_dist"], "pub_every": pub_div},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{args.steps} windows of {args.workload} after {args.warmup} "
                                   f"warm-up; {desc}; cv2 threads = {cores}; the reference's own "
                                   "node cannot be built here (needs ROS/OpenCV C++/Eigen)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stage_seconds": timers,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
class _CudaArray:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<i4",
                                         "data": (ptr, False), "version": 3, "strides": None}


def init_nccl(local, p2p_peer=None):
    """NCCL prints its version banner to stdout when the first communicator comes up; stdout
    must carry exactly one JSON line, so fd 1 points at stderr until that has happened."""
    import torch
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        warm = torch.zeros(1, device=torch.device("cuda", local))
        dist.all_reduce(warm)
        if p2p_peer is not None:        # the send/recv communicator comes up lazily as well
            if dist.get_rank() < p2p_peer:
                dist.recv(warm, src=p2p_peer)
            else:
                dist.send(warm, dst=p2p_peer)
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)


def main_split(args):
    """--split-lr, 2 ranks: ONE stereo stream with the right camera's SAE / time surface /
    pyramid on rank 1 and everything else on rank 0 (SURVEY.md 8e row 2, 8d config 3 "then 2
    GPUs with L/R split"); the right image block crosses NVLink once per window (NCCL
    send/recv).  Strong scaling of one stream; rank 0 also runs the same windows on one handle
    and reports that throughput and whether the results are identical."""
    import torch
    import torch.distributed as dist
    from esvio_b200 import frontend

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world != 2:
        raise SystemExit("bench.py --split-lr needs exactly 2 ranks (torchrun --nproc-per-node 2)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the front-end has no CPU fallback")
    torch.cuda.set_device(local)
    init_nccl(local, p2p_peer=1 - rank)
    dev = torch.device("cuda", local)
    w, cfg, pub_div = workload_cfg(args.workload)
    n_per_cam = int(round(w["rate"] / synth.WINDOWS_PER_SEC))
    cfg = dict(cfg, device_id=local, max_events_per_window=max(n_per_cam + 64, 1024))
    K, Wm = args.steps, args.warmup
    wins = gen_windows(w, 0, K + Wm)                 # both ranks see the same stereo stream
    fe = frontend.EventFrontEnd(cfg)
    sp = shard.LeftRightSplit(fe, rank)
    mine = [frontend._Ev(frontend.DeviceEvents(fe, (L, R)[rank])) for L, R, _ in wins]
    n_ev_local = float(sum(len((L, R)[rank][0]) for L, R, _ in wins[Wm:]))
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    DEPTH = 3

    def run(k0, n, sink):
        waited = 0
        for k in range(k0, k0 + n):
            sp.step(wins[k][2], mine[k], k % pub_div == 0)
            if rank == shard.LEFT_RANK and k - k0 >= DEPTH - 1:
                sink.append(sp.wait(unpack=False))
                waited += 1
        while rank == shard.LEFT_RANK and waited < n:
            sink.append(sp.wait(unpack=False))
            waited += 1

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    counts = []
    run(0, Wm, counts)
    flush.fill_(1)
    cur = torch.cuda.current_stream()
    tstream = torch.cuda.ExternalStream(fe.stream(), device=dev) if rank == shard.LEFT_RANK else cur
    barrier()
    launches0 = fe.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tstream)
    run(Wm, K, counts)
    e1.record(tstream)
    barrier()
    launches = fe.kernel_launches() - launches0
    value, ms = shard.aggregate_throughput(n_ev_local, e0.elapsed_time(e1))
    nl = torch.tensor([launches], dtype=torch.int64, device=dev)
    dist.all_reduce(nl)
    img_bytes = fe.split_right_buffer()[1] if rank == shard.LEFT_RANK else 0

    one = None
    if rank == shard.LEFT_RANK:       # the same windows on one handle, same pipelining
        fe1 = frontend.EventFrontEnd(cfg)
        dw = [(frontend._Ev(frontend.DeviceEvents(fe1, L)), frontend._Ev(frontend.DeviceEvents(fe1, R)), t)
              for L, R, t in wins]
        ext1 = torch.cuda.ExternalStream(fe1.stream(), device=dev)
        ref_counts = []

        def run1(k0, n):
            waited = 0
            for k in range(k0, k0 + n):
                fe1.submit(dw[k][2], dw[k][0], dw[k][1], k % pub_div == 0)
                if k - k0 >= DEPTH - 1:
                    ref_counts.append(fe1.wait(unpack=False))
                    waited += 1
            while waited < n:
                ref_counts.append(fe1.wait(unpack=False))
                waited += 1

        run1(0, Wm)
        flush.fill_(2)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(ext1)
        run1(Wm, K)
        f1.record(ext1)
        torch.cuda.synchronize()
        ms1 = f0.elapsed_time(f1)
        n_ev = float(sum(len(L[0]) + len(R[0]) for L, R, _ in wins[Wm:]))
        # last window in full, every window by its feature counts
        a, b = fe._unpack(), fe1._unpack()
        same = ref_counts == counts and all(np.array_equal(a[k], b[k]) for k in a if k != "stats")
        one = {"value": n_ev / (ms1 * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms1 / K,
               "identical_results": bool(same)}
        fe1.close()
    dist.barrier()
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 2, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64 SAE / u8 time surface / f32 LK", "data": "synthetic",
            "config": {"workload": args.workload, "parallelism": "lr_split: rank 0 left camera + "
                       "tracking, rank 1 right camera SAE/time surface/pyramid",
                       "inputs": "resident in HBM on the rank that consumes them",
                       "windows_in_flight": DEPTH},
            "exchange": {"what": "right pyramid block, NCCL send/recv per window",
                         "bytes_per_step": int(img_bytes)},
            "gpu_launches": int(nl.item()), "one_gpu_same_run": one}))
    fe.close()
    dist.destroy_process_group()


def frames_leg(width, height, device_id, n_frames=48, cpu_frames=16):
    """Extra record at N = 1: FeatureTracker::trackImage on synthetic stereo
    frames of the workload's resolution through esvio_fe_track_image_submit / _wait, three frames
    in flight, host frames copied inside the timed region; every 2nd frame is a publish frame
    (Image_setMask + goodFeaturesToTrack).  CPU beside it: the oracle's trackImage with OpenCV
    LK, 1 thread.  Never raises: a failure is reported in the record."""
    rec = {"what": "trackImage, stereo frames/s", "width": width, "height": height}
    try:
        import torch
        from esvio_b200 import frontend
        mc, md = (150, 10) if width < 600 else (175, 40)   # config/esvio, config/esvio_DSEC
        cfg = synth.default_config(width, height, max_cnt=mc, min_dist=md)
        rec.update(max_cnt_img=mc, min_dist_img=md, frames=n_frames, pub_every=2)
        warm = 4
        frames = synth.stereo_frame_sequence(width, height, n_frames + warm)
        fe = frontend.EventFrontEnd(dict(cfg, device_id=device_id, max_events_per_window=1024))
        # host frames in pinned memory (esvio_fe_host_alloc), as a driver's DMA buffers would be
        import ctypes
        lib, pinned = frontend._capi.lib(), []

        def pin(a):
            ptr = lib.esvio_fe_host_alloc(a.nbytes)
            if not ptr:
                return a
            pinned.append(ptr)
            v = np.frombuffer((ctypes.c_uint8 * a.nbytes).from_address(ptr), np.uint8).reshape(a.shape)
            v[:] = a
            return v

        frames = [(pin(l), pin(r)) for l, r in frames]
        rec["host_frames"] = "pinned" if pinned else "pageable"
        ext = torch.cuda.ExternalStream(fe.stream(), device=torch.device("cuda", device_id))
        for k in range(warm):
            fe.track_image(1.0 + k / 20.0, frames[k][0], frames[k][1], k % 2 == 0)
        torch.cuda.synchronize()
        launches0 = fe.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(ext)
        waited, last = 0, (0, 0)
        for k in range(warm, warm + n_frames):
            fe.submit_image(1.0 + k / 20.0, frames[k][0], frames[k][1], k % 2 == 0)
            if k - warm >= 2:
                last = fe.wait(unpack=False)
                waited += 1
        while waited < n_frames:
            last = fe.wait(unpack=False)
            waited += 1
        e1.record(ext)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        rec.update(value=n_frames / (ms * 1e-3), unit="stereo frames/s", ms_per_frame=ms / n_frames,
                   wall_ms_per_frame=wall_ms / n_frames,
                   gpu_launches=int(fe.kernel_launches() - launches0),
                   h2d_bytes_per_frame=2 * width * height,
                   tracks_last_frame={"left": int(last[0]), "right": int(last[1])})
        fe.close()
        frames = [(np.array(l), np.array(r)) for l, r in frames]   # the CPU leg reads copies
        for ptr in pinned:
            lib.esvio_fe_host_free(ptr)
        try:
            from oracle import oracle as ora   # CPU leg of the bench: the checker as a baseline
            ora.build()
            trk = ora.OracleTracker(cfg, use_cv2=True, cv2_threads=1)
            n_cpu = min(cpu_frames, n_frames)
            for k in range(2):
                trk.track_image(1.0 + k / 20.0, frames[k][0], frames[k][1], k % 2 == 0)
            t0 = time.perf_counter()
            for k in range(2, 2 + n_cpu):
                trk.track_image(1.0 + k / 20.0, frames[k][0], frames[k][1], k % 2 == 0)
            dt = time.perf_counter() - t0
            rec["cpu_baseline"] = {"value": n_cpu / dt, "unit": "stereo frames/s", "cores": 1,
                                   "kind": "port", "sample": f"{n_cpu} frames, oracle trackImage "
                                   "(C goodFeaturesToTrack restatement + cv2 LK)"}
        except Exception as e:  # noqa: BLE001
            rec["cpu_baseline"] = {"error": repr(e)[:200]}
    except Exception as e:  # noqa: BLE001
        rec["error"] = repr(e)[:300]
    return rec


def main_ours(args):
    import torch
    import torch.distributed as dist
    from esvio_b200 import frontend

    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the front-end has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        init_nccl(local)
    dev = torch.device("cuda", local)

    w, cfg, pub_div = workload_cfg(args.workload)
    n_per_cam = int(round(w["rate"] / synth.WINDOWS_PER_SEC))
    cfg = dict(cfg, device_id=local, max_events_per_window=max(n_per_cam + 64, 1024))
    K, Wm = args.steps, args.warmup
    wins = gen_windows(w, rank, K + Wm)
    ev_per_step = sum(len(L[0]) + len(R[0]) for L, R, _ in wins[Wm:]) / max(K, 1)

    # ---------------- device-resident leg: `value` ----------------
    fe = frontend.EventFrontEnd(cfg)
    ext = torch.cuda.ExternalStream(fe.stream(), device=dev)
    dwins = [(frontend._Ev(frontend.DeviceEvents(fe, L)), frontend._Ev(frontend.DeviceEvents(fe, R)), t)
             for L, R, t in wins]
    rptr, rbytes = fe.result_device_ptr()
    res_t = torch.as_tensor(_CudaArray(rptr, rbytes), device=dev)
    gathered = torch.empty((world, res_t.numel()), dtype=torch.int32, device=dev) if world > 1 else None
    input_bytes = 13 * ev_per_step * K
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step_submit(k):
        l, r, t = dwins[k]
        fe.submit(t, l, r, k % pub_div == 0)
        if world > 1:
            with torch.cuda.stream(ext):   # stream-ordered behind this window's finalize
                shard.all_gather_tracks(res_t, gathered)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(Wm):
        step_submit(k)
        fe.wait(unpack=False)
    DEPTH = 3  # windows in flight: event stage | temporal stage | stereo stage

    def run_pipelined(k0, n, collect=None):
        """n windows from k0, three in flight; returns the (n_left, n_right) of the last one."""
        last = (0, 0)
        for k in range(k0, min(k0 + DEPTH - 1, k0 + n)):
            step_submit(k)
        waited = 0
        for k in range(k0 + DEPTH - 1, k0 + n):
            step_submit(k)
            last = fe.wait(unpack=False)
            waited += 1
            if collect is not None:
                collect(fe.stage_ms())
        while waited < n:
            last = fe.wait(unpack=False)
            waited += 1
            if collect is not None:
                collect(fe.stage_ms())
        return last

    flush.fill_(1)  # evict the uploaded windows: every timed step streams its events from HBM
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    launches0 = fe.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    n_left_last, n_right_last = run_pipelined(Wm, K)
    e1.record(ext)
    barrier()
    launches = fe.kernel_launches() - launches0
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    # second pipelined pass over the head of the same windows with the per-stage CUDA events
    # on: stage_ms and the roofline kernel's launch duration.  The events sit between the
    # kernels and switch off their programmatic launch overlap, so they stay out of the pass
    # that yields `value`.
    fe.reset()
    Kp = min(K, 100)
    for k in range(Wm):
        step_submit(k)
        fe.wait(unpack=False)
    flush.fill_(4)
    fe.set_profiling(True)
    stage_sum = np.zeros(len(frontend._capi.STAGE_NAMES))

    def collect(d):
        nonlocal stage_sum
        stage_sum += np.fromiter(d.values(), float)

    barrier()
    run_pipelined(Wm, Kp, collect)
    barrier()
    fe.set_profiling(False)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([ev_per_step * K, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms_max = float(t_ms.item())
    value = float(tot[0].item()) / (ms_max * 1e-3) / 1e6
    gpu_launches = int(tot[1].item())
    fe.close()

    # ---------------- end-to-end leg: host buffers through the synchronous C-ABI call ----------
    fe2 = frontend.EventFrontEnd(cfg)
    ext2 = torch.cuda.ExternalStream(fe2.stream(), device=dev)
    pwins = [(frontend._Ev(frontend.PinnedEvents(L)), frontend._Ev(frontend.PinnedEvents(R)), t)
             for L, R, t in wins]
    def e2e_submit(k):
        l, r, t = pwins[k]
        fe2.submit(t, l, r, k % pub_div == 0)   # H2D of the events is enqueued inside
        if world > 1:
            with torch.cuda.stream(ext2):
                shard.all_gather_tracks(res2_t, gathered)

    res2_t = torch.as_tensor(_CudaArray(*fe2.result_device_ptr()), device=dev)
    for k in range(Wm):
        e2e_submit(k)
        fe2.wait(unpack=False)
    flush.fill_(2)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(ext2)
    checksum = 0
    for k in range(Wm, min(Wm + DEPTH - 1, Wm + K)):
        e2e_submit(k)
    waited = 0
    for k in range(Wm + DEPTH - 1, Wm + K):
        e2e_submit(k)                     # window k's copies + event stage overlap earlier LK
        nl, nr = fe2.wait(unpack=False)   # D2H of the track records + host sync
        checksum += nl + nr
        waited += 1
    while waited < K:
        nl, nr = fe2.wait(unpack=False)
        checksum += nl + nr
        waited += 1
    g1.record(ext2)
    barrier()
    e2e_ms = g0.elapsed_time(g1)
    # the same windows once more through the synchronous drop-in call, for reference
    sync_ms = None
    k1_alone = []
    if world == 1:
        fe4 = frontend.EventFrontEnd(cfg)
        n_sync = min(K, 60)
        for k in range(Wm):
            l, r, t = pwins[k]
            fe4.track_raw(t, l, r, k % pub_div == 0)
        torch.cuda.synchronize()
        fe4.set_profiling(True)   # one window at a time: the kernels run without neighbours
        t0 = time.perf_counter()
        for k in range(Wm, Wm + n_sync):
            l, r, t = pwins[k]
            fe4.track_raw(t, l, r, k % pub_div == 0)
            k1_alone.append(fe4.stage_ms()["sae_update_ts"])
        sync_ms = (time.perf_counter() - t0) * 1e3 / n_sync
        fe4.close()
    # ---------------- batched leg (N = 1): S streams of this workload in one group ----------
    # SURVEY.md 8d caveat: one window of one stream moves 6-25 MB per k_sae_update_ts launch,
    # i.e. 1-4 us at the HBM peak -- the roofline of that kernel only means something when a
    # launch covers several streams (BASELINE configs[4], esvio_fe_group_*).
    batched = None
    if world == 1 and args.batch_streams > 1:
        S = args.batch_streams
        Kb = min(K, args.batch_steps)
        nb_w = Wm + Kb
        grp = frontend.EventFrontEndGroup(cfg, S)
        m0 = grp.member(0)
        bw = []
        for i in range(S):
            ws = wins[:nb_w] if i == 0 else gen_windows(w, 100 + i, nb_w)
            bw.append([(frontend._Ev(frontend.DeviceEvents(m0, L)), frontend._Ev(frontend.DeviceEvents(m0, R)), t,
                        len(L[0]) + len(R[0])) for L, R, t in ws])
        def gsub(k):
            grp.submit([bw[i][k][2] for i in range(S)], [bw[i][k][0] for i in range(S)],
                       [bw[i][k][1] for i in range(S)], [k % pub_div == 0] * S)
        for k in range(Wm):
            gsub(k)
            grp.wait(unpack=False)
        flush.fill_(3)
        torch.cuda.synchronize()
        k1_ms = []
        t0 = time.perf_counter()
        gsub(Wm)
        if Kb > 1:
            gsub(Wm + 1)
        done = 0
        for k in range(Wm + 2, Wm + Kb):
            gsub(k)
            grp.wait(unpack=False)
            done += 1
            k1_ms.append(grp.sae_ts_ms())
        while done < Kb:
            grp.wait(unpack=False)
            done += 1
            k1_ms.append(grp.sae_ts_ms())
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        n_ev_b = sum(bw[i][k][3] for i in range(S) for k in range(Wm, Wm + Kb))
        k1 = float(np.mean(k1_ms))
        alg = S * 2 * 17 * w["width"] * w["height"] + 45 * n_ev_b / Kb
        peak_b, _ = peaks()
        batched = {"streams": S, "steps": Kb, "value": n_ev_b / (wall_ms * 1e-3) / 1e6, "unit": UNIT,
                   "ms_per_step": wall_ms / Kb, "timing": "host wall clock around submit/wait with "
                   "synchronize on both sides, events device-resident, 3 windows in flight",
                   "gpu_launches": grp.kernel_launches(),
                   "roofline": {"bound": "hbm", "kernel": "k_sae_update_ts", "kernel_ms": k1,
                                "algorithmic_bytes_per_launch": int(alg),
                                "achieved": alg / (k1 * 1e-3) / 1e9, "peak": peak_b, "unit": "GB/s",
                                "frac": alg / (k1 * 1e-3) / 1e9 / peak_b,
                                "note": "one launch covers the 2*S cameras of the group; CUDA events "
                                        "around the launch on the group's event-stage stream, other "
                                        "stages of other windows share the SMs"}}
        grp.close()
    e2e_t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = float(tot[0].item()) / (float(e2e_t.item()) * 1e-3) / 1e6
    h2d = int(round(13 * ev_per_step))
    d2h = int(rbytes)

    # ---------------- CPU baseline (rank 0, N = 1 only) + parity of the first windows ---------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_s = min(len(wins), args.cpu_windows)
        n_ev, sec, outs, kind, desc, _ = run_cpu(workload_cfg(args.workload)[1], pub_div,
                                                 wins[:n_s], min(Wm, 3), 1)
        cpu = {"value": n_ev / sec / 1e6, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"first {n_s} windows of {args.workload} ({n_ev} events timed), single "
                         f"thread like the reference's worker (stereo_event_tracker_node.cpp:366); {desc}"}
        fe3 = frontend.EventFrontEnd(cfg)
        sq, cnt, mx, hor = 0.0, 0, 0.0, 0
        for k in range(min(n_s, 8)):
            L, R, t = wins[k]
            g = fe3.track(t, L, R, k % pub_div == 0)
            o = outs[k]
            if not (np.array_equal(g["id"], o["id"]) and np.array_equal(g["id_right"], o["id_right"])):
                break
            hor = k + 1
            for a, b in ((g["u"], o["u"]), (g["v"], o["v"]), (g["ru"], o["ru"]), (g["rv"], o["rv"])):
                if len(a):
                    d = np.abs(a - b)
                    sq += float((d ** 2).sum())
                    cnt += len(d)
                    mx = max(mx, float(d.max()))
        parity = {"tracked_px_rmse_vs_ref": (sq / max(cnt, 1)) ** 0.5, "max_px": mx,
                  "windows_with_identical_ids": hor, "coords_compared": cnt,
                  "ref": "oracle with OpenCV LK" if "cv2" in desc else "oracle C LK"}
        fe3.close()
    fe2.close()

    if rank == 0:
        peak, peak_src = peaks()
        W_, H_ = w["width"], w["height"]
        names = frontend._capi.STAGE_NAMES
        stage_ms = dict(zip(names, (stage_sum / max(Kp, 1)).tolist()))
        k1_ms = stage_ms["sae_update_ts"]
        alg_bytes = 2 * 17 * W_ * H_ + 45 * ev_per_step   # 17*W*H per camera + 45 B/event
        achieved = alg_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0
        gpu_ms = sum(v for k_, v in stage_ms.items() if k_ not in ("h2d", "d2h"))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms_max / max(K, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "width": W_, "height": H_,
                       "events_per_window_per_camera": n_per_cam, "max_cnt": w["max_cnt"],
                       "min_dist": w["min_dist"], "pub_every": pub_div,
                       "streams_per_gpu": 1, "parallelism": f"{world} independent stereo streams"
                       + (", NCCL all-gather of track records per window" if world > 1 else ""),
                       "l2": "each window's events are read once from HBM: all windows are "
                             "uploaded, then L2 is flushed with a 512 MiB write before the timed "
                             "region; the SAE state (the path's persistent working set) stays "
                             "resident by design",
                       "timed_input_bytes": int(input_bytes)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": float(e2e_t.item()) / max(K, 1),
                    "api": "esvio_fe_track_submit / esvio_fe_track_wait on pinned host SoA "
                           "buffers, three windows in flight (H2D + event stage of window k+2 | "
                           "temporal LK + selection of k+1 | stereo LK of k)",
                    "sync_call_ms_per_step": sync_ms},
            "gpu_launches": gpu_launches,
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "k_sae_update_ts", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.workload),
                         "algorithmic_bytes_per_launch": int(alg_bytes),
                         "kernel_ms": k1_ms, "share_of_step": k1_ms / gpu_ms if gpu_ms else None,
                         "kernel_ms_alone": float(np.mean(k1_alone)) if k1_alone else None,
                         "frac_alone": (alg_bytes / (float(np.mean(k1_alone)) * 1e-3) / 1e9 / peak)
                         if k1_alone else None,
                         "peak_source": peak_src,
                         "note": "one launch covers both cameras of one window; achieved = "
                                 "algorithmic bytes (SURVEY.md 8d: 17*W*H per camera + 45 B/event) / "
                                 "CUDA-event time of the launch in a second pipelined pass over the same windows "
                                 "with the per-stage events on "
                                 "(the LK / selection kernels of two other windows share the SMs; "
                                 "`kernel_ms_alone` / `frac_alone`: the same launch in the synchronous "
                                 "call, nothing else on the GPU); the SAE state is L2-resident between "
                                 "windows, `traffic` is the DRAM traffic of one launch under ncu "
                                 "(caches flushed); LK stages are latency-bound and reported by time"},
            "stage_ms": stage_ms,
            "tracks_last_window": {"left": int(n_left_last), "right": int(n_right_last)},
        }
        if batched is not None:
            line["batched"] = batched
        if world == 1 and not args.no_frames:
            line["frames"] = frames_leg(W_, H_, local)
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if parity is not None:
            line["parity"] = parity
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(synth.WORKLOADS))
    ap.add_argument("--cpu-windows", type=int, default=150)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--batch-streams", type=int, default=8,
                    help="streams of the extra batched leg at N=1 (esvio_fe_group); 1 disables it")
    ap.add_argument("--batch-steps", type=int, default=60)
    ap.add_argument("--no-frames", action="store_true", help="skip the trackImage record")
    ap.add_argument("--split-lr", action="store_true",
                    help="2 ranks: one stereo stream split by camera (SURVEY.md 8e row 2)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        main_reference(args)
    elif args.split_lr:
        main_split(args)
    else:
        main_ours(args)


if __name__ == "__main__":
    main()
