"""Replay harness: a raw stereo event recording -> fixed-rate windows -> left/right pairing ->
handle_stereo_event -> feature clouds, all through the GPU tracker.

  python -m esvio_b200.replay [--workload stereo_davis346_1mevs] [--windows 30] [--npz rec.npz]
                              [--config config/esvio_DSEC]
  python -m esvio_b200.replay --frames [--windows 40] [--npz frames.npz]

`--config` takes a configuration set of the reference (a directory with es*io.yaml and the two
event calibrations, or the yaml itself): resolution, MAX_CNT, MIN_DIST, FREQ, EQUALIZE, ... and
the pinhole models come from it exactly as readParameters_event reads them (esvio_b200/config.py).

`--npz` replays a recording with arrays lx, ly, lt, lp, rx, ry, rt, rp (time-ascending,
seconds); without it the synthetic stream of the named workload is used.  `--frames` replays
stereo frames instead (image pairing step -> handle_stereo_image -> trackImage, SURVEY.md 8f rank
4): arrays left[k], right[k] (uint8, H x W), stamps_left, stamps_right, or synthetic frames at
20 Hz of the workload's resolution.  Prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import time

import numpy as np

from . import config as config_sets
from . import node, synth


def synthetic_recording(name: str, windows: int):
    w = synth.WORKLOADS[name]
    s = synth.StereoEventStream(w["width"], w["height"], w["rate"], mono=w["mono"])
    L = [s.window(k, 0) for k in range(windows)]
    R = [s.window(k, 1) for k in range(windows)]
    cat = lambda ws: tuple(np.concatenate([q[i] for q in ws]) for i in range(4))
    return w, cat(L), cat(R)


def frame_messages(args):
    """(width, height, freq, left ImageMsg list, right ImageMsg list)"""
    w = synth.WORKLOADS[args.workload]
    if args.npz:
        z = np.load(args.npz)
        L, R = z["left"], z["right"]
        tl, tr = z["stamps_left"], z["stamps_right"]
        H, W = L.shape[1:]
    else:
        W, H = w["width"], w["height"]
        seq = synth.stereo_frame_sequence(W, H, args.windows)
        L, R = [f[0] for f in seq], [f[1] for f in seq]
        tl = 1.7e9 + 0.05 * np.arange(len(L))
        tr = tl + 0.002
    lm = [node.ImageMsg(t, img) for t, img in zip(tl, L)]
    rm = [node.ImageMsg(t, img) for t, img in zip(tr, R)]
    return W, H, w["freq"], lm, rm


def replay_frames(args, frontend):
    W, H, freq, lm, rm = frame_messages(args)
    mc, md = (150, 10) if W < 600 else (175, 40)     # config/esvio, config/esvio_DSEC
    cfg = synth.default_config(W, H, max_cnt=mc, min_dist=md)
    cfg["max_events_per_window"] = 1024
    ft = frontend.FeatureTracker(cfg)
    nd = node.StereoImageNode(ft, freq)
    t0 = time.perf_counter()
    clouds, dropped = node.replay_images(nd, lm, rm)
    dt = time.perf_counter() - t0
    print(json.dumps({
        "frames": [len(lm), len(rm)], "width": W, "height": H, "frames_tracked": nd.windows_tracked,
        "clouds_published": len(clouds),
        "rows_last_cloud": int(len(clouds[-1].rows)) if clouds else 0, "queue_overwrites": dropped,
        "restarts": nd.restarts, "tracking_s": dt,
        "frames_per_s_sync_call": nd.windows_tracked / max(dt, 1e-9)}))
    ft.fe.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="stereo_davis346_1mevs", choices=sorted(synth.WORKLOADS))
    ap.add_argument("--windows", type=int, default=30)
    ap.add_argument("--npz", default=None)
    ap.add_argument("--frequency", type=float, default=30.0, help="re-windowing rate (EventMessageEditor: 30)")
    ap.add_argument("--frames", action="store_true", help="replay stereo frames through trackImage")
    ap.add_argument("--config", default=None,
                    help="reference configuration set: directory with es*io.yaml, or the yaml")
    args = ap.parse_args()
    from . import frontend  # needs libesvio_fe.so and a B200: there is no CPU path

    if args.frames:
        return replay_frames(args, frontend)
    w, left, right = synthetic_recording(args.workload, args.windows)
    if args.npz:
        z = np.load(args.npz)
        left = tuple(z[k] for k in ("lx", "ly", "lt", "lp"))
        right = tuple(z[k] for k in ("rx", "ry", "rt", "rp"))
    cfg = synth.default_config(w["width"], w["height"], max_cnt=w["max_cnt"], min_dist=w["min_dist"],
                               use_ransac=1)
    freq, do_mc = w["freq"], False
    if args.config:
        import os
        f = config_sets.find_config(args.config) if os.path.isdir(args.config) else args.config
        cfg, nparams = config_sets.read_parameters_event(f)
        if (cfg["width"], cfg["height"]) != (w["width"], w["height"]) and not args.npz:
            raise SystemExit(f"--config is {cfg['width']}x{cfg['height']}, the synthetic workload "
                             f"{args.workload} is {w['width']}x{w['height']}: pick a matching --workload")
        freq, do_mc = nparams["freq"], bool(cfg["do_motion_correction"])
    cfg["max_events_per_window"] = max(1 << 16, int(2.5 * w["rate"] / args.frequency))
    ft = frontend.FeatureTracker(cfg)
    nd = node.StereoEventNode(ft, freq, do_motion_correction=do_mc)
    t0 = time.perf_counter()
    lm, rm = node.window_stream(left, args.frequency), node.window_stream(right, args.frequency)
    t1 = time.perf_counter()
    clouds, dropped = node.replay(nd, lm, rm)
    t2 = time.perf_counter()
    n_ev = sum(len(m) for m in lm) + sum(len(m) for m in rm)
    print(json.dumps({
        "workload": args.workload, "config": args.config, "max_cnt": cfg["max_cnt"], "min_dist": cfg["min_dist"],
        "freq": freq, "messages": [len(lm), len(rm)], "events": n_ev,
        "windows_tracked": nd.windows_tracked, "clouds_published": len(clouds),
        "rows_last_cloud": int(len(clouds[-1].rows)) if clouds else 0, "queue_overwrites": dropped,
        "restarts": nd.restarts, "windowing_s": t1 - t0, "tracking_s": t2 - t1,
        "mevents_per_s_sync_call": n_ev / max(t2 - t1, 1e-9) / 1e6}))
    ft.fe.close()


if __name__ == "__main__":
    main()
