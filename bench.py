#!/usr/bin/env python
"""bench.py -- Mevents/s through the event front-end (SAE + time surface + Arc* + temporal and
stereo LK) on synthetic stereo event streams (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (libesvio_fe.so)
  python bench.py --impl reference --gpus N --steps K ...  the CPU path on the host cores

A "step" is one window (1/30 s of stream time) of one stereo pair: both cameras' events go
through createSAE/time surface, the left camera through corner detection, then temporal and
stereo LK (FeatureTracker::trackEvent, feature_tracker/src/feature_tracker.cpp:340-603).

Workload: N = 1 runs BASELINE.json configs[2] (stereo 640x480 @ 5 Mev/s per camera, the
configuration north_star quotes its target on); N > 1 runs configs[4]'s stream (640x480 @
10 Mev/s per camera), one independent stereo stream per rank (weak scaling, SURVEY.md 8e
"independent streams"), the packed track records all-gathered over NCCL on publish windows.
One JSON line is printed by rank 0.  `config` is identical in both arms.

Extra records on the N = 1 line: `rigid_scene` (the same workload on a scene with ONE epipolar
geometry, both arms: the survey's scene of 64 independent movers makes the CPU arm's F-RANSAC
run its full iteration budget), `sync` (the synchronous drop-in call), `secondary`
(configs[1]) and `scale_base` (configs[4]'s stream on one GPU: the denominator of the scaling
curve), `batched` (a group of streams on the one GPU, where the roofline kernel is measured),
`frames` (FeatureTracker::trackImage).

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
WORKLOAD_N1 = "stereo_vga_5mevs"      # BASELINE.json configs[2]: the north-star configuration
WORKLOAD_NX = "stereo_vga_10mevs"     # configs[4]: one of the 8 streams, per rank
WORKLOAD_SECONDARY = "stereo_davis346_1mevs"  # configs[1]
DEPTH = 3  # windows in flight; set to esvio_fe_pipeline_depth() once the library is loaded


def env_int(k, d):
    return int(os.environ.get(k, d))


# stdout must carry exactly ONE JSON line, but NCCL (with NCCL_DEBUG set) and others write to
# fd 1 whenever a communicator comes up: fd 1 points at stderr for the whole run, and the line
# goes to the saved descriptor.
_STDOUT_FD = None


def claim_stdout():
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: str):
    sys.stdout.flush()
    if _STDOUT_FD is None:
        print(line)
    else:
        os.write(_STDOUT_FD, (line + "\n").encode())


def default_workload(world):
    return WORKLOAD_N1 if world == 1 else WORKLOAD_NX


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_sae_update_ts launch, from the
    committed `ncu --set full` capture (profiles/r2_k1_ncu_full.json, else round 1's)."""
    for name in ("r2_k1_ncu_full.json", "r1_k1_ncu_full.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return int(json.load(f)[key]["traffic_bytes_per_launch"])
        except Exception:
            continue
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
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
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
    pub_div = int(round(synth.WINDOWS_PER_SEC / w["freq"]))  # freq 10 -> every 3rd window
    return w, cfg, pub_div


def config_of(name, scene="survey"):
    """The `config` object of the JSON line -- the same in both arms."""
    w, _, pub_div = workload_cfg(name)
    return {"workload": name, "width": w["width"], "height": w["height"],
            "events_per_window_per_camera": int(round(w["rate"] / synth.WINDOWS_PER_SEC)),
            "max_cnt": w["max_cnt"], "min_dist": w["min_dist"], "pub_every": pub_div,
            "scene": scene}


def gen_windows(w, stream, n, scene="survey"):
    s = synth.StereoEventStream(w["width"], w["height"], w["rate"], stream=stream, mono=w["mono"],
                                rigid=(scene == "rigid"))
    return [s.stereo_window(k) for k in range(n)]


def n_events(wins):
    return float(sum(len(L[0]) + len(R[0]) for L, R, _ in wins))


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


def run_cpu(cfg, pub_div, wins, warmup, threads, barrier=None):
    t, kind, desc = cpu_tracker(cfg, threads)
    outs = []
    n_ev = 0
    t_total = 0.0
    for k, (L, R, tc) in enumerate(wins):
        if k == warmup and barrier is not None:
            barrier.wait()
        t0 = time.perf_counter()
        o = t.track(tc, L, R, k % pub_div == 0)
        dt = time.perf_counter() - t0
        if k >= warmup:
            t_total += dt
            n_ev += len(L[0]) + len(R[0])
        outs.append(o)
    return n_ev, t_total, outs, kind, desc, t.timers()


def _ref_stream_worker(workload, scene, stream, steps, warmup, threads, barrier, q):
    """One CPU stream of the reference arm at N > 1 (its own process, `threads` OpenCV threads)."""
    try:
        w, cfg, pub_div = workload_cfg(workload)
        wins = gen_windows(w, stream, steps + warmup, scene)
        n_ev, sec, _, kind, desc, timers = run_cpu(cfg, pub_div, wins, warmup, threads, barrier)
        q.put((stream, n_ev, sec, kind, desc, timers))
    except Exception as e:  # noqa: BLE001
        try:
            barrier.abort()
        except Exception:
            pass
        q.put((stream, 0, 0.0, "error", repr(e)[:200], {}))


def reference_measure(workload, scene, n_streams, steps, warmup, cores):
    """(Mevents/s, seconds, kind, desc, timers, threads per stream): n_streams CPU streams side
    by side on `cores` host threads; value = all events / the slowest stream's timed seconds."""
    if n_streams == 1:
        w, cfg, pub_div = workload_cfg(workload)
        wins = gen_windows(w, 0, steps + warmup, scene)
        n_ev, sec, _, kind, desc, timers = run_cpu(cfg, pub_div, wins, warmup, cores)
        return n_ev / sec / 1e6, sec, kind, desc, timers, cores
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    threads = max(1, cores // n_streams)
    barrier = ctx.Barrier(n_streams)
    q = ctx.Queue()
    procs = [ctx.Process(target=_ref_stream_worker,
                         args=(workload, scene, s, steps, warmup, threads, barrier, q))
             for s in range(n_streams)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    bad = [r for r in res if r[3] == "error"]
    if bad:
        raise RuntimeError("reference stream failed: " + bad[0][4])
    n_ev = sum(r[1] for r in res)
    sec = max(r[2] for r in res)
    return n_ev / sec / 1e6, sec, res[0][3], res[0][4], res[0][5], threads


def reference_code_record(workload, n_windows=12, warmup=3, threads=None):
    """The reference's OWN FeatureTracker::trackEvent (feature_tracker.cpp + event_detector.cc
    compiled unmodified, oracle/_ref/libesvio_ref_ft.so) with its OpenCV calls served by REAL
    OpenCV (cv2, all host threads) through the stand-in headers: the closest thing to the
    reference node's tracker that can run here (Eigen and cv::Mat storage are stand-ins).  It is
    SLOWER than the oracle + cv2 arm above (same OpenCV, reference-authored C++ around it), so
    the headline CPU arm stays the faster one and this record is reported next to it."""
    from oracle import ref_tracker
    L = ref_tracker.load()
    if L is None:
        return {"unavailable": "oracle/_ref/libesvio_ref_ft.so not built"}
    from esvio_b200 import synth
    from oracle import oracle as ora
    real = ora.have_cv2()
    if real:
        ref_tracker.use_real_opencv(L, True, threads=threads)
    w, cfg, pub_div = workload_cfg(workload)
    s = synth.StereoEventStream(w["width"], w["height"], w["rate"], mono=w["mono"])
    cfg = synth.default_config(w["width"], w["height"], max_cnt=w["max_cnt"], min_dist=w["min_dist"])
    wins = [(s.window(k, 0), s.window(k, 1)) for k in range(warmup + n_windows)]
    rt = ref_tracker.RefTracker(L, cfg)
    try:
        ev = 0
        for k, (a, b) in enumerate(wins):
            if k == warmup:
                t0 = time.perf_counter()
            rt.track(float(a[2][-1]), a, b, k % pub_div == 0)
            if k >= warmup:
                ev += len(a[0]) + len(b[0])
        sec = time.perf_counter() - t0
    finally:
        rt.close()
        if real:
            ref_tracker.use_real_opencv(L, False)
    return {"value": ev / sec / 1e6, "unit": UNIT, "cores": threads or (os.cpu_count() or 1), "kind": "reference",
            "ms_per_step": 1e3 * sec / n_windows,
            "sample": f"{n_windows} windows of {workload} after {warmup} warm-up through the reference's own "
                      "FeatureTracker::trackEvent (unmodified feature_tracker.cpp + event_detector.cc, "
                      "oracle/_ref); OpenCV calls = " + ("real OpenCV (cv2) through the stand-in headers"
                                                       if real else "the oracle's scalar C restatements")}


def main_reference(args):
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return
    n_streams = max(1, args.gpus)
    workload = args.workload or default_workload(n_streams)
    cores = os.cpu_count() or 1
    val, sec, kind, desc, timers, threads = reference_measure(workload, "survey", n_streams,
                                                              args.steps, args.warmup, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_of(workload),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{n_streams} stream(s) x {args.steps} windows of {workload} after "
                                   f"{args.warmup} warm-up; {desc}; {threads} OpenCV threads per stream"
                                   + (", one process per stream" if n_streams > 1 else "")
                                   + "; the reference's own node cannot be built here (needs "
                                     "ROS/OpenCV C++/Eigen); its feature_tracker.cpp + event_detector.cc, "
                                     "compiled unmodified against stand-in headers (oracle/_ref), pin this "
                                     "oracle bit for bit and are timed as `reference_code`"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stage_seconds": timers,
    }
    if n_streams == 1:
        try:
            line["reference_code"] = reference_code_record(workload)
        except Exception as e:  # informational: must not take the line down
            line["reference_code"] = {"error": repr(e)}
    if n_streams == 1 and not args.no_rigid:
        rv, rsec, _, _, rtimers, _ = reference_measure(workload, "rigid", 1, args.steps,
                                                       args.warmup, cores)
        line["rigid_scene"] = {"config": config_of(workload, "rigid"), "value": rv, "unit": UNIT,
                               "e2e": {"value": rv, "unit": UNIT},
                               "ms_per_step": 1e3 * rsec / max(args.steps, 1),
                               "stage_seconds": rtimers}
    emit(json.dumps(line))


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
class _CudaArray:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<i4",
                                         "data": (ptr, False), "version": 3, "strides": None}


def init_nccl(local):
    import torch
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warm = torch.zeros(1, device=torch.device("cuda", local))
    dist.all_reduce(warm)
    torch.cuda.synchronize()


def run_pipelined(fe, wins, k0, n, pub_div, gather=False, collect=None):
    """n windows from k0 through submit/wait, a pipeline depth in flight; `wins[k]` = (left, right,
    t) as esvio_events wrappers.  `gather`: the replica mode's collective -- on publish windows
    (the reference publishes nothing on the others, stereo_event_tracker_node.cpp:268) the packed
    track block is all-gathered over NCCL by the library itself (esvio_fe_allgather_tracks), on
    a stream of its own.  Returns (n_left, n_right) of the last window and their sum over all."""
    last, checksum, waited = (0, 0), 0, 0
    for k in range(k0, k0 + n):
        l, r, t = wins[k]
        pub = k % pub_div == 0
        fe.submit(t, l, r, pub)
        if gather and pub:
            fe.allgather_tracks()
        if k - k0 >= DEPTH - 1:
            last = fe.wait(unpack=False)
            checksum += last[0] + last[1]
            waited += 1
            if collect is not None:
                collect(fe.stage_ms())
    while waited < n:
        last = fe.wait(unpack=False)
        checksum += last[0] + last[1]
        waited += 1
        if collect is not None:
            collect(fe.stage_ms())
    return last, checksum


class StreamBench:
    """One stereo stream on this rank's GPU: the device-resident leg (`value`), the host-buffer
    leg through submit/wait (`e2e`), the synchronous drop-in call and the per-stage profile."""

    def __init__(self, torch, dev, local, workload, scene, stream_id, K, Wm, world=1):
        from esvio_b200 import frontend
        self.torch, self.dev, self.fr = torch, dev, frontend
        self.w, cfg, self.pub_div = workload_cfg(workload)
        self.n_per_cam = int(round(self.w["rate"] / synth.WINDOWS_PER_SEC))
        self.cfg = dict(cfg, device_id=local, max_events_per_window=max(self.n_per_cam + 64, 1024))
        self.K, self.Wm, self.world = K, Wm, world
        self.rank = stream_id if world > 1 else 0
        self.wins = gen_windows(self.w, stream_id, K + Wm, scene)
        self.ev_timed = n_events(self.wins[Wm:])
        self.ev_per_step = self.ev_timed / max(K, 1)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def _timed(self, fe, wins, gather, flush, fill):
        torch = self.torch
        ext = torch.cuda.ExternalStream(fe.stream(), device=self.dev)
        for k in range(self.Wm):
            l, r, t = wins[k]
            pub = k % self.pub_div == 0
            fe.submit(t, l, r, pub)
            if gather and pub:
                fe.allgather_tracks()
            fe.wait(unpack=False)
        flush.fill_(fill)   # evict the uploaded windows: every timed step streams its events from HBM
        self.barrier()
        launches0 = fe.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        last, checksum = run_pipelined(fe, wins, self.Wm, self.K, self.pub_div, gather)
        if gather:   # the timed region ends when the last all-gather has
            ext.wait_stream(torch.cuda.ExternalStream(fe.gathered_tracks()[2], device=self.dev))
        e1.record(ext)
        self.barrier()
        if gather:   # every rank's block of the last publish window arrived intact
            ptr, nbytes, _ = fe.gathered_tracks()
            blocks = shard.device_bytes(ptr, nbytes * self.world).cpu().numpy().view(np.int32)
            blocks = blocks.reshape(self.world, -1)
            M = self.cfg["max_cnt"]
            self.gather_ok = bool(all(0 < b[0] <= M and 0 <= b[1] <= M for b in blocks)
                                  and len({int(b[9]) for b in blocks}) >= 1)
        return e0.elapsed_time(e1), fe.kernel_launches() - launches0, last, checksum

    def device_leg(self, flush, comm_id=None, profile=False):
        fr = self.fr
        fe = fr.EventFrontEnd(self.cfg)
        gather = comm_id is not None
        if gather:
            fe.comm_init(comm_id, self.rank, self.world)
        held = [(fr.DeviceEvents(fe, L), fr.DeviceEvents(fe, R)) for L, R, _ in self.wins]
        dw = [(fr._Ev(a), fr._Ev(b), w[2]) for (a, b), w in zip(held, self.wins)]
        ms, launches, last, _ = self._timed(fe, dw, gather, flush, 1)
        out = {"ms": ms, "launches": launches, "last": last, "result_bytes": fe.result_device_ptr()[1]}
        if profile:
            # second pipelined pass over the same windows with the per-stage CUDA events on:
            # stage_ms and the SAE kernel's launch duration.  The events sit between the kernels
            # and switch off their programmatic launch overlap, so they stay out of the pass
            # that yields `value`.
            fe.reset()
            for k in range(self.Wm):
                l, r, t = dw[k]
                fe.submit(t, l, r, k % self.pub_div == 0)
                fe.wait(unpack=False)
            flush.fill_(4)
            fe.set_profiling(True)
            acc = np.zeros(len(fr._capi.STAGE_NAMES))

            def collect(d):
                acc[:] += np.fromiter(d.values(), float)

            self.barrier()
            run_pipelined(fe, dw, self.Wm, self.K, self.pub_div, False, collect)
            self.barrier()
            fe.set_profiling(False)
            out["stage_ms"] = dict(zip(fr._capi.STAGE_NAMES, (acc / max(self.K, 1)).tolist()))
        for a, b in held:
            a.free()
            b.free()
        fe.close()
        return out

    def host_leg(self, flush, comm_id=None, sync=False):
        """e2e: pinned host SoA buffers (one block per window) through esvio_fe_track_submit / _wait (H2D of the events
        and D2H of the track records inside the timed region); `sync`: the same windows once
        more through the synchronous esvio_fe_track, host wall clock."""
        fr = self.fr
        fe = fr.EventFrontEnd(self.cfg)
        gather = comm_id is not None
        if gather:
            fe.comm_init(comm_id, self.rank, self.world)
        # both cameras of a window in one pinned block (esvio_fe_soa_layout_stereo): one transfer
        blocks = [fr.PinnedStereoEvents(L, R) for L, R, _ in self.wins]
        pw = [(fr._Ev(b.left), fr._Ev(b.right), w[2]) for b, w in zip(blocks, self.wins)]
        ms, launches, last, checksum = self._timed(fe, pw, gather, flush, 2)
        out = {"ms": ms, "launches": launches, "last": last, "checksum": checksum}
        fe.close()
        if sync:
            fe = fr.EventFrontEnd(self.cfg)
            for k in range(self.Wm):
                l, r, t = pw[k]
                fe.track_raw(t, l, r, k % self.pub_div == 0)
            self.torch.cuda.synchronize()
            # every 4th window runs with the per-stage CUDA events on (k_sae_update_ts "alone": one
            # window at a time, no neighbours) and is left out of the wall-clock figure: the
            # events and their read-back cost the call 10-20 us.  Publish windows (every
            # pub_div-th) keep their share among the timed ones.
            k1, wall = [], {True: [], False: []}
            for k in range(self.Wm, self.Wm + self.K):
                l, r, t = pw[k]
                prof = k % 4 == 3
                fe.set_profiling(prof)
                t0 = time.perf_counter()
                fe.track_raw(t, l, r, k % self.pub_div == 0)
                dt = time.perf_counter() - t0
                if prof:
                    k1.append(fe.stage_ms()["sae_update_ts"])
                else:
                    wall[k % self.pub_div == 0].append(dt)
            share = 1.0 / self.pub_div   # of publish windows in the stream
            mean = lambda v: float(np.mean(v)) if v else 0.0   # noqa: E731
            out["sync_ms_per_step"] = (share * mean(wall[True]) + (1.0 - share) * mean(wall[False])) * 1e3
            out["sync_ms_publish"], out["sync_ms_other"] = mean(wall[True]) * 1e3, mean(wall[False]) * 1e3
            out["k1_alone_ms"] = float(np.mean(k1))
            fe.close()
        return out

    def dropin_leg(self, scene="survey", stream_id=0):
        """The call exactly as the reference node would make it: esvio_fe_track on the 16-byte
        dvs_msgs::Event records (AoS: u16 x, u16 y, u32 sec, u32 nsec, u8 polarity) of a
        std::vector in ordinary, pageable host memory (stereo_event_tracker_node.cpp:193 holds the
        events of both cameras that way); host wall clock, one window at a time."""
        fr = self.fr
        s = synth.StereoEventStream(self.w["width"], self.w["height"], self.w["rate"], stream=stream_id,
                                    mono=self.w["mono"], rigid=(scene == "rigid"))
        wins = []
        for k in range(self.K + self.Wm):
            recs = []
            for cam in (0, 1):
                x, y, t, p, sec, nsec = s.window(k, cam)
                recs.append(fr._Ev(synth.to_aos(x, y, sec, nsec, p)))
            wins.append((recs[0], recs[1], self.wins[k][2]))
        fe = fr.EventFrontEnd(self.cfg)
        for k in range(self.Wm):
            fe.track_raw(wins[k][2], wins[k][0], wins[k][1], k % self.pub_div == 0)
        self.torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(self.Wm, self.Wm + self.K):
            n = fe.track_raw(wins[k][2], wins[k][0], wins[k][1], k % self.pub_div == 0)
        ms = (time.perf_counter() - t0) * 1e3 / self.K
        fe.close()
        return {"value": self.ev_per_step / (ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms,
                "h2d_bytes_per_step": int(round(16 * self.ev_per_step)),
                "tracks_last_window": {"left": int(n[0]), "right": int(n[1])},
                "api": "esvio_fe_track (synchronous) on 16-byte dvs_msgs::Event records (AoS) in ordinary "
                       "pageable host memory -- the buffers the reference node's callback holds "
                       "(stereo_event_tracker_node.cpp:128-142,193); host wall clock"}

    def mev(self, ms):
        return self.ev_timed / (ms * 1e-3) / 1e6


def quick_record(torch, dev, local, workload, scene, K, Wm, flush, sync=False):
    """value + e2e (+ sync) of one more workload / scene on this GPU, as a record."""
    sb = StreamBench(torch, dev, local, workload, scene, 0, K, Wm)
    d = sb.device_leg(flush)
    h = sb.host_leg(flush, sync=sync)
    rec = {"config": config_of(workload, scene), "value": sb.mev(d["ms"]), "unit": UNIT,
           "ms_per_step": d["ms"] / K,
           "e2e": {"value": sb.mev(h["ms"]), "unit": UNIT, "ms_per_step": h["ms"] / K,
                   "h2d_bytes_per_step": int(round(13 * sb.ev_per_step)),
                   "d2h_bytes_per_step": int(d["result_bytes"])},
           "tracks_last_window": {"left": int(d["last"][0]), "right": int(d["last"][1])}}
    if sync:
        rec["sync"] = {"value": sb.ev_per_step / (h["sync_ms_per_step"] * 1e-3) / 1e6, "unit": UNIT,
                       "ms_per_step": h["sync_ms_per_step"], "ms_publish_window": h["sync_ms_publish"],
                       "ms_other_window": h["sync_ms_other"]}
    return rec


def batched_leg(torch, dev, local, workload, S, K, Wm, flush):
    """S streams of the workload in one esvio_fe_group on this GPU (BASELINE configs[4] on one
    GPU): ONE k_sae_update_ts launch per window covers the 2S cameras.  This is where the
    roofline of that kernel is measured (SURVEY.md 8d caveat: one window of one stream moves
    6-25 MB per launch, i.e. 1-4 us at the HBM peak): CUDA events around the launch on the
    group's event-stage stream while the tracking stages of two other windows x S streams share
    the SMs; whole-group throughput by CUDA events as well."""
    from esvio_b200 import frontend
    w, cfg, pub_div = workload_cfg(workload)
    n_per_cam = int(round(w["rate"] / synth.WINDOWS_PER_SEC))
    cfg = dict(cfg, device_id=local, max_events_per_window=max(n_per_cam + 64, 1024))
    grp = frontend.EventFrontEndGroup(cfg, S)
    m0 = grp.member(0)
    n_alone = 6   # synchronous windows after the pipelined pass: the same launch with the GPU to itself
    nb_w = Wm + K + n_alone
    bw, n_ev, held = [], 0.0, []
    for i in range(S):
        ws = gen_windows(w, 100 + i, nb_w)
        n_ev += n_events(ws[Wm:Wm + K])
        row = []
        for L, R, t in ws:
            a, b = frontend.DeviceEvents(m0, L), frontend.DeviceEvents(m0, R)
            held += [a, b]
            row.append((frontend._Ev(a), frontend._Ev(b), t))
        bw.append(row)

    def gsub(k):
        grp.submit([bw[i][k][2] for i in range(S)], [bw[i][k][0] for i in range(S)],
                   [bw[i][k][1] for i in range(S)], [k % pub_div == 0] * S)

    for k in range(Wm):
        gsub(k)
        grp.wait(unpack=False)
    flush.fill_(3)
    torch.cuda.synchronize()
    ext = torch.cuda.ExternalStream(m0.stream(), device=dev)   # member 0's result stream
    k1_ms = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(ext)
    waited = 0
    for k in range(Wm, Wm + K):
        gsub(k)
        if k - Wm >= DEPTH - 1:
            grp.wait(unpack=False)
            waited += 1
            k1_ms.append(grp.sae_ts_ms())
    while waited < K:
        grp.wait(unpack=False)
        waited += 1
        k1_ms.append(grp.sae_ts_ms())
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    k1 = float(np.mean(k1_ms))
    k1_alone = []
    for k in range(Wm + K, nb_w):
        gsub(k)
        grp.wait(unpack=False)
        k1_alone.append(grp.sae_ts_ms())
    k1_alone = float(np.mean(k1_alone[2:])) if len(k1_alone) > 2 else None
    alg = S * 2 * 17 * w["width"] * w["height"] + 45 * n_ev / K
    peak, peak_src = peaks()
    rec = {"streams": S, "steps": K, "config": config_of(workload),
           "value": n_ev / (wall_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": wall_ms / K,
           "timing": "host wall clock around submit/wait with synchronize on both sides, events "
                     f"device-resident, {DEPTH} windows in flight",
           "gpu_launches": grp.kernel_launches()}
    roof = {"bound": "hbm", "kernel": "k_sae_update_ts", "achieved": alg / (k1 * 1e-3) / 1e9,
            "peak": peak, "unit": "GB/s", "frac": alg / (k1 * 1e-3) / 1e9 / peak,
            "traffic": ncu_traffic(f"{workload}_x{S}"),
            "algorithmic_bytes_per_launch": int(alg), "kernel_ms": k1,
            "share_of_step": k1 / (wall_ms / K), "peak_source": peak_src,
            "kernel_ms_alone": k1_alone,
            "frac_alone": (alg / (k1_alone * 1e-3) / 1e9 / peak) if k1_alone else None,
            "launch_covers": f"{2 * S} cameras = {S} stereo streams of {workload} in one esvio_fe_group",
            "note": "achieved = algorithmic bytes (SURVEY.md 8d: 17*W*H per camera + 45 B/event, "
                    "summed over the cameras of the launch) / CUDA-event time of the launch, "
                    "measured inside the pipeline (the LK / selection kernels of two other "
                    "windows x S streams share the SMs: k_lk holds 64 registers x 256 threads x 4 "
                    "CTAs = the whole register file of an SM, so this launch gets 4 of its 14 CTAs "
                    "per SM there); *_alone = the same launch in synchronous group windows right "
                    "after; the group's SAE state (S x 19.7 MB at 640x480) exceeds L2, so the "
                    "launch streams from HBM (traffic: whole 16 KB tiles, profiles/r2_k1_l2_policy.txt)"}
    for a in held:
        a.free()
    grp.close()
    return rec, roof


def frames_leg(width, height, device_id, n_frames=48, cpu_frames=16):
    """Extra record at N = 1: FeatureTracker::trackImage on synthetic stereo
    frames of the workload's resolution through esvio_fe_track_image_submit / _wait, a pipeline
    depth of frames in flight, host frames copied inside the timed region; every 2nd frame is a publish frame
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
            if k - warm >= frontend.pipeline_depth() - 1:
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


def parity_record(frontend, cfg, pub_div, wins, outs, desc):
    """Tracked-px RMSE of the first windows against the CPU arm's outputs (lock-step windows)."""
    fe = frontend.EventFrontEnd(cfg)
    sq, cnt, mx, hor = 0.0, 0, 0.0, 0
    for k in range(min(len(outs), 8)):
        L, R, t = wins[k]
        g = fe.track(t, L, R, k % pub_div == 0)
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
    fe.close()
    return {"tracked_px_rmse_vs_ref": (sq / max(cnt, 1)) ** 0.5, "max_px": mx,
            "windows_with_identical_ids": hor, "coords_compared": cnt,
            "ref": "oracle with OpenCV LK" if "cv2" in desc else "oracle C LK",
            "all_windows": "tests/test_gpu_parity.py::test_teacher_forced_every_window compares every "
                           "window of the run, restarted from the reference state each window"}


def bind_to_gpu_cpus(torch, local):
    """CPU affinity of this rank = the CPUs NVML lists as local to its GPU, so that the pinned event
    buffers are first-touched on the GPU's NUMA node and the host->device copies of N ranks do not
    share the socket interconnect (N = 4 moved 110 GB/s in aggregate on one box and 184 GB/s on
    another with unbound ranks).  Applied only if at least 4 of the process's CPUs remain.
    Returns a description for the `run` record."""
    try:
        import pynvml
        before = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(local).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        want = cpus & before
        if len(want) >= 4 and want != before:
            os.sched_setaffinity(0, want)
            return f"gpu-local ({len(want)} of {len(before)} CPUs)"
        return f"unchanged ({len(before)} CPUs, {len(want)} of them gpu-local)"
    except Exception as e:  # no NVML, no permission, CPUs outside the container's set ...
        return f"unchanged ({type(e).__name__})"


def main_ours(args):
    import torch
    import torch.distributed as dist
    from esvio_b200 import frontend

    global DEPTH
    DEPTH = frontend.pipeline_depth()
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the front-end has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    affinity = bind_to_gpu_cpus(torch, local)
    if world > 1:
        init_nccl(local)
    dev = torch.device("cuda", local)
    workload = args.workload or default_workload(world)
    K, Wm = args.steps, args.warmup
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

    sb = StreamBench(torch, dev, local, workload, "survey", rank, K, Wm, world)
    comm_ids = [None, None]
    if world > 1:   # one NCCL communicator per handle: rank 0 makes the ids, torch ships them
        comm_ids = [frontend.nccl_unique_id(), frontend.nccl_unique_id()] if rank == 0 else [None, None]
        dist.broadcast_object_list(comm_ids, src=0)
    clocks = ClockSampler(local)
    clocks.start()
    d = sb.device_leg(flush, comm_ids[0], profile=True)
    h = sb.host_leg(flush, comm_ids[1], sync=(world == 1))
    clk = clocks.stop()

    red = torch.tensor([d["ms"], h["ms"]], dtype=torch.float64, device=dev)
    tot = torch.tensor([sb.ev_timed, float(d["launches"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms_max, e2e_ms = float(red[0].item()), float(red[1].item())
    ev_all = float(tot[0].item())
    value = ev_all / (ms_max * 1e-3) / 1e6
    e2e_value = ev_all / (e2e_ms * 1e-3) / 1e6

    # N > 1: the same stream alone on rank 0's GPU while the other ranks idle -- the one-GPU
    # figure of THIS workload in THIS run (the N = 1 line runs configs[2])
    one_gpu = None
    if world > 1:
        if rank == 0:
            sb1 = StreamBench(torch, dev, local, workload, "survey", 0, K, Wm, 1)
            one_gpu = {"value": sb1.mev(sb1.device_leg(flush)["ms"]), "unit": UNIT,
                       "what": "rank 0's stream alone, no collective, same run"}
        dist.barrier()

    # N = 2: SURVEY.md 8e row 2 on the same two GPUs -- ONE stream of configs[2] split by camera
    split_lr = None
    if world == 2 and not args.no_split:
        try:
            split_lr = measure_split(torch, dist, frontend, rank, local, dev, WORKLOAD_N1, K, Wm, flush)
        except Exception as e:  # noqa: BLE001
            split_lr = {"error": repr(e)[:300]}

    # N = 4: SURVEY.md 8e row 3 on the same GPUs -- ONE stream of configs[3] sharded by time window
    time_shard = None
    if world == 4 and not args.no_split:
        try:
            time_shard = measure_time_shard(torch, dist, frontend, rank, world, local, dev,
                                            "stereo_vga_20mevs_burst", K, Wm, flush)
        except Exception as e:  # noqa: BLE001
            time_shard = {"error": repr(e)[:300]}

    extra = {}
    if rank == 0 and world == 1:
        if not args.no_rigid:
            extra["rigid_scene"] = quick_record(torch, dev, local, workload, "rigid", K, Wm, flush, sync=True)
        if not args.no_secondary:
            if workload != WORKLOAD_SECONDARY:
                extra["secondary"] = quick_record(torch, dev, local, WORKLOAD_SECONDARY, "survey", K, Wm, flush)
            if workload != WORKLOAD_NX:
                extra["scale_base"] = quick_record(torch, dev, local, WORKLOAD_NX, "survey", K, Wm, flush)
        roof_group = None
        if args.batch_streams > 1:
            try:
                extra["batched"], roof_group = batched_leg(torch, dev, local, workload, args.batch_streams,
                                                           min(K, args.batch_steps), Wm, flush)
            except Exception as e:  # noqa: BLE001
                extra["batched"] = {"error": repr(e)[:300]}
        if not args.no_frames:
            extra["frames"] = frames_leg(sb.w["width"], sb.w["height"], local)
        if not args.no_cpu:
            _, cfg0, pub_div = workload_cfg(workload)
            n_s = min(len(sb.wins), args.cpu_windows)
            n_ev, sec, outs, kind, desc, _ = run_cpu(cfg0, pub_div, sb.wins[:n_s], min(Wm, 3), 1)
            extra["cpu_baseline"] = {
                "value": n_ev / sec / 1e6, "unit": UNIT, "cores": 1, "kind": kind,
                "sample": f"first {n_s} windows of {workload} ({n_ev} events timed), single thread "
                          f"like the reference's worker (stereo_event_tracker_node.cpp:366); {desc}"}
            extra["parity"] = parity_record(frontend, sb.cfg, pub_div, sb.wins, outs, desc)

    if rank == 0:
        peak, peak_src = peaks()
        W_, H_ = sb.w["width"], sb.w["height"]
        stage_ms = d["stage_ms"]
        k1_ms = stage_ms["sae_update_ts"]
        alg_bytes = 2 * 17 * W_ * H_ + 45 * sb.ev_per_step   # 17*W*H per camera + 45 B/event
        achieved = alg_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0
        gpu_ms = sum(v for k_, v in stage_ms.items() if k_ not in ("h2d", "d2h"))
        roof_single = {
            "bound": "hbm", "kernel": "k_sae_update_ts", "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(workload),
            "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": k1_ms,
            "share_of_step": k1_ms / gpu_ms if gpu_ms else None,
            "kernel_ms_alone": h.get("k1_alone_ms"),
            "frac_alone": (alg_bytes / (h["k1_alone_ms"] * 1e-3) / 1e9 / peak) if h.get("k1_alone_ms") else None,
            "peak_source": peak_src, "launch_covers": "the 2 cameras of one window of one stream",
            "note": "one window of one stream moves 1-4 us worth of bytes at the HBM peak, less "
                    "than a kernel launch costs (SURVEY.md 8d caveat); see `roofline` for the "
                    "launch that covers a group of streams"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms_max / max(K, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(workload),
            "run": {"streams_per_gpu": 1, "parallelism": f"{world} independent stereo streams"
                    + (", NCCL all-gather of the packed track records on publish windows, "
                       "enqueued by the library (esvio_fe_allgather_tracks) on a stream of its own"
                       if world > 1 else ""),
                    "windows_in_flight": DEPTH, "cpu_affinity": affinity,
                    "l2": "each window's events are read once from HBM: all windows are uploaded, "
                          "then L2 is flushed with a 512 MiB write before the timed region; the SAE "
                          "state (the path's persistent working set) stays resident by design",
                    "timed_input_bytes": int(13 * sb.ev_timed)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(round(13 * sb.ev_per_step)),
                    "d2h_bytes_per_step": int(d["result_bytes"]), "ms_per_step": e2e_ms / max(K, 1),
                    "api": "esvio_fe_track_submit / esvio_fe_track_wait on pinned host SoA "
                           f"buffers, {DEPTH} windows in flight (the windows' copies and kernels "
                           "overlap on several streams)",
                    "sync_call_ms_per_step": h.get("sync_ms_per_step")},
            "gpu_launches": int(tot[1].item()),
            "clocks": clk,
            "stage_ms": stage_ms,
            "tracks_last_window": {"left": int(d["last"][0]), "right": int(d["last"][1])},
        }
        if h.get("sync_ms_per_step"):
            line["sync"] = {"value": sb.ev_per_step / (h["sync_ms_per_step"] * 1e-3) / 1e6, "unit": UNIT,
                            "ms_per_step": h["sync_ms_per_step"],
                            "ms_publish_window": h.get("sync_ms_publish"),
                            "ms_other_window": h.get("sync_ms_other"),
                            "api": "esvio_fe_track: the synchronous call the reference node makes "
                                   "(stereo_event_tracker_node.cpp:193), pinned host buffers, host "
                                   "wall clock"}
        if world == 1:
            try:
                line["dropin"] = sb.dropin_leg("survey", rank)
            except Exception as e:  # a secondary record must not take the line down
                line["dropin"] = {"error": repr(e)}
        if world == 1 and extra.get("batched") and "error" not in extra["batched"] and roof_group:
            line["roofline"] = roof_group
            line["roofline_single_stream"] = roof_single
        else:
            line["roofline"] = roof_single
        if world > 1:
            line["run"]["all_gather_delivered_every_rank"] = getattr(sb, "gather_ok", None)
        if one_gpu is not None:
            line["one_gpu_same_workload"] = one_gpu
        if split_lr is not None:
            line["split_lr"] = split_lr
        if time_shard is not None:
            line["time_shard"] = time_shard
        line.update(extra)
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_split(torch, dist, frontend, rank, local, dev, workload, K, Wm, flush):
    """ONE stereo stream with the right camera's SAE / time surface / pyramid on rank 1 and
    everything else on rank 0 (SURVEY.md 8e row 2, 8d config 3 "then 2 GPUs with L/R split"); the
    right image block crosses NVLink once per window (NCCL send/recv).  Strong scaling of one
    stream; rank 0 also runs the same windows on one handle and reports that throughput and
    whether the results are identical.  Returns the record on rank 0, None on rank 1."""
    warm = torch.zeros(1, device=dev)      # the send/recv communicator comes up lazily
    if rank == 0:
        dist.recv(warm, src=1)
    else:
        dist.send(warm, dst=0)
    torch.cuda.synchronize()
    w, cfg, pub_div = workload_cfg(workload)
    n_per_cam = int(round(w["rate"] / synth.WINDOWS_PER_SEC))
    cfg = dict(cfg, device_id=local, max_events_per_window=max(n_per_cam + 64, 1024))
    wins = gen_windows(w, 0, K + Wm)                 # both ranks see the same stereo stream
    fe = frontend.EventFrontEnd(cfg)
    sp = shard.LeftRightSplit(fe, rank)
    held = [frontend.DeviceEvents(fe, (L, R)[rank]) for L, R, _ in wins]
    mine = [frontend._Ev(h) for h in held]
    n_ev_local = float(sum(len((L, R)[rank][0]) for L, R, _ in wins[Wm:]))

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
        held1 = [(frontend.DeviceEvents(fe1, L), frontend.DeviceEvents(fe1, R)) for L, R, _ in wins]
        dw = [(frontend._Ev(a), frontend._Ev(b), wn[2]) for (a, b), wn in zip(held1, wins)]
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
        n_ev = n_events(wins[Wm:])
        # last window in full, every window by its feature counts
        a, b = fe._unpack(), fe1._unpack()
        same = ref_counts == counts and all(np.array_equal(a[k], b[k]) for k in a if k != "stats")
        one = {"value": n_ev / (ms1 * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms1 / K,
               "identical_results": bool(same)}
        for x, y in held1:
            x.free()
            y.free()
        fe1.close()
    dist.barrier()
    for h in held:
        h.free()
    fe.close()
    if rank != 0:
        return None
    return {"value": value, "unit": UNIT, "n_gpus": 2, "steps": K, "warmup": Wm, "ms_per_step": ms / K,
            "scaling": "strong", "config": config_of(workload),
            "run": {"parallelism": "lr_split: rank 0 left camera + tracking, rank 1 right camera "
                                   "SAE/time surface/pyramid",
                    "inputs": "resident in HBM on the rank that consumes them",
                    "windows_in_flight": DEPTH},
            "exchange": {"what": "right pyramid block, NCCL send/recv per window",
                         "bytes_per_step": int(img_bytes)},
            "gpu_launches": int(nl.item()), "one_gpu_same_run": one}


def measure_time_shard(torch, dist, frontend, rank, world, local, dev, workload, K, Wm, flush):
    """ONE stereo stream whose SAE / time-surface / corner stages are sharded by time window over
    `world` GPUs (BASELINE configs[3]): rank r replays window r of every
    round of `world` windows with the carry-in protocol of shard.TimeShardRank (two all-gathers of
    the state planes and one of the windows' images + corner candidates per round, NCCL), rank 0
    runs the serial track chain.  Strong scaling of one stream; rank 0 also runs the same
    windows through the ordinary pipelined path on its own and reports that throughput and
    whether the results are identical.  Returns the record on rank 0, None elsewhere."""
    w, cfg, pub_div = workload_cfg(workload)
    n_per_cam = int(round(w["rate"] / synth.WINDOWS_PER_SEC))
    cfg = dict(cfg, device_id=local, max_events_per_window=max(n_per_cam + 64, 1024))
    R = world
    n_rounds_w, n_rounds = -(-Wm // R), -(-K // R)
    n_win = (n_rounds_w + n_rounds) * R
    k_timed0 = n_rounds_w * R
    wins = gen_windows(w, 0, n_win)
    fe = frontend.EventFrontEnd(cfg)
    stream = torch.cuda.Stream(device=dev)
    rk = shard.TimeShardRank(fe, rank, R, stream)
    held = {k: (frontend.DeviceEvents(fe, wins[k][0]), frontend.DeviceEvents(fe, wins[k][1]))
            for k in range(rank, n_win, R)}
    mine = {k: (frontend._Ev(a), frontend._Ev(b)) for k, (a, b) in held.items()}
    empty = frontend._Ev(None)
    planes = [torch.empty((R, rk.nd), dtype=torch.float64, device=dev) for _ in range(2)]
    prods = torch.empty((R, rk.prod_bytes), dtype=torch.uint8, device=dev)
    tracker = tr = None
    if rank == 0:
        tracker = frontend.EventFrontEnd(cfg)
        tr = shard.TimeShardTracker(tracker, rk)
    counts, pending = [], [0]

    def one_round(rnd):
        k = rnd * R + rank
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(planes[0].view(-1), rk.phase_a(wins[k][2], *mine[k]))
            dist.all_gather_into_tensor(planes[1].view(-1), rk.phase_c(planes[0]))
            dist.all_gather_into_tensor(prods.view(-1), rk.phase_d(planes[1], k % pub_div == 0, empty, empty))
        if rank == 0:
            for r in range(R):
                kk = rnd * R + r
                if pending[0] >= DEPTH:
                    counts.append(tracker.wait(unpack=False))
                    pending[0] -= 1
                tr.submit(prods[r], wins[kk][2], len(wins[kk][0][0]), kk % pub_div == 0)
                pending[0] += 1
            # the products buffer is rewritten by the next round's all-gather: the tracker's
            # copies out of it are ordered on `stream`, like the all-gather

    def drain():
        while rank == 0 and pending[0] > 0:
            counts.append(tracker.wait(unpack=False))
            pending[0] -= 1

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for rnd in range(n_rounds_w):
        one_round(rnd)
    drain()
    flush.fill_(5)
    barrier()
    tstream = torch.cuda.ExternalStream(tracker.stream(), device=dev) if rank == 0 else stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tstream)
    for rnd in range(n_rounds_w, n_rounds_w + n_rounds):
        one_round(rnd)
    drain()
    e1.record(tstream)
    barrier()
    n_ev = n_events(wins[k_timed0:])
    value, ms = shard.aggregate_throughput(n_ev if rank == 0 else 0.0, e0.elapsed_time(e1))
    nl = torch.tensor([fe.kernel_launches() + (tracker.kernel_launches() if tracker else 0)],
                      dtype=torch.int64, device=dev)
    dist.all_reduce(nl)
    one = None
    if rank == 0:
        fe1 = frontend.EventFrontEnd(cfg)
        held1 = [(frontend.DeviceEvents(fe1, L), frontend.DeviceEvents(fe1, Rr)) for L, Rr, _ in wins]
        dw = [(frontend._Ev(a), frontend._Ev(b), wn[2]) for (a, b), wn in zip(held1, wins)]
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

        run1(0, k_timed0)
        flush.fill_(6)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(ext1)
        run1(k_timed0, n_win - k_timed0)
        f1.record(ext1)
        torch.cuda.synchronize()
        ms1 = f0.elapsed_time(f1)
        a, b = tracker._unpack(), fe1._unpack()
        same = ref_counts == counts and all(np.array_equal(a[k], b[k]) for k in a if k != "stats")
        one = {"value": n_ev / (ms1 * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms1 / (n_win - k_timed0),
               "identical_results": bool(same)}
        for x, y in held1:
            x.free()
            y.free()
        fe1.close()
        tracker.close()
    dist.barrier()
    for a, b in held.values():
        a.free()
        b.free()
    fe.close()
    if rank != 0:
        return None
    return {"value": value, "unit": UNIT, "n_gpus": world, "steps": n_win - k_timed0, "warmup": k_timed0,
            "ms_per_step": ms / (n_win - k_timed0), "scaling": "strong", "config": config_of(workload),
            "run": {"parallelism": f"time-window shard: rank r replays window r of every round of {R} "
                                   "windows (SAE update, time surface, pyramids, Arc* candidates); "
                                   "rank 0 also runs the serial track chain",
                    "inputs": "resident in HBM on the rank that consumes them",
                    "windows_in_flight": DEPTH},
            "exchange": {"what": "per round: all-gather of the last-event planes, all-gather of the "
                                 "accepted-time planes, all-gather of the windows' image pyramids + "
                                 "corner candidate lists (NCCL)",
                         "bytes_per_round_per_rank": int(2 * rk.nd * 8 + rk.prod_bytes)},
            "gpu_launches": int(nl.item()), "one_gpu_same_run": one}


def main_split(args):
    """--split-lr, 2 ranks: only the left/right split measurement (the N = 2 line of the default
    bench carries the same record as `split_lr`)."""
    import torch
    import torch.distributed as dist
    from esvio_b200 import frontend

    global DEPTH
    DEPTH = frontend.pipeline_depth()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.split_lr and world != 2:
        raise SystemExit("bench.py --split-lr needs exactly 2 ranks (torchrun --nproc-per-node 2)")
    if args.time_shard and world < 2:
        raise SystemExit("bench.py --time-shard needs >= 2 ranks (torchrun --nproc-per-node N)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the front-end has no CPU fallback")
    torch.cuda.set_device(local)
    init_nccl(local)
    dev = torch.device("cuda", local)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    if args.time_shard:
        rec = measure_time_shard(torch, dist, frontend, rank, world, local, dev,
                                 args.workload or "stereo_vga_20mevs_burst", args.steps, args.warmup, flush)
    else:
        rec = measure_split(torch, dist, frontend, rank, local, dev, args.workload or WORKLOAD_N1,
                            args.steps, args.warmup, flush)
    if rank == 0:
        line = {"metric": METRIC, "higher_is_better": True, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic"}
        line.update(rec)
        emit(json.dumps(line))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(synth.WORKLOADS),
                    help=f"default: {WORKLOAD_N1} at N = 1, {WORKLOAD_NX} at N > 1")
    ap.add_argument("--cpu-windows", type=int, default=40)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-rigid", action="store_true", help="skip the rigid-scene record")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[1] / configs[4] records")
    ap.add_argument("--batch-streams", type=int, default=8,
                    help="streams of the batched leg at N=1 (esvio_fe_group); 1 disables it")
    ap.add_argument("--batch-steps", type=int, default=30)
    ap.add_argument("--no-frames", action="store_true", help="skip the trackImage record")
    ap.add_argument("--no-split", action="store_true", help="N = 2: skip the left/right split record")
    ap.add_argument("--time-shard", action="store_true",
                    help="N ranks: only the time-window shard measurement (SURVEY.md 8e row 3)")
    ap.add_argument("--split-lr", action="store_true",
                    help="2 ranks: one stereo stream split by camera (SURVEY.md 8e row 2)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    claim_stdout()
    if args.impl == "reference":
        main_reference(args)
    elif args.split_lr or args.time_shard:
        main_split(args)
    else:
        main_ours(args)


if __name__ == "__main__":
    main()
