"""A few windows through every kernel path, small enough to run under compute-sanitizer:
  compute-sanitizer --tool memcheck  python scratch/sanitize.py
  compute-sanitizer --tool racecheck python scratch/sanitize.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esvio_b200 import frontend, synth

W, H = 346, 260
s = synth.StereoEventStream(W, H, 3.0e5)          # 10 000 events per camera and window
wins = [s.stereo_window(k) for k in range(4)]
motion = dict(state_v=(1.0, 0.5, 0.2), v_pre=(0.9, 0.45, 0.25), accel=(4.0, 3.0, 2.0), omega=(0.5, -0.3, 0.8))
for kw in (dict(), dict(equalize=1, median_blur_kernel_size=1), dict(do_motion_correction=1)):
    cfg = synth.default_config(W, H, use_ransac=1, max_cnt=60, max_events_per_window=1 << 14, **kw)
    fe = frontend.EventFrontEnd(cfg)
    for k, (L, R, t) in enumerate(wins):
        m = dict(motion, t1=float(L[2][-1])) if kw.get("do_motion_correction") else None
        o = fe.track(t, L, R, k % 2 == 0, motion=m)
    print("single", kw, len(o["id"]), len(o["id_right"]))
    # pipelined submit/wait
    fe.reset()
    fe.submit(wins[0][2], wins[0][0], wins[0][1], True)
    fe.submit(wins[1][2], wins[1][0], wins[1][1], False)
    fe.submit(wins[2][2], wins[2][0], wins[2][1], True)
    for _ in range(3):
        fe.wait()
    fe.close()
cfg = synth.default_config(W, H, use_ransac=1, max_cnt=60, max_events_per_window=1 << 14)
g = frontend.EventFrontEndGroup(cfg, 2)
for k, (L, R, t) in enumerate(wins):
    out = g.track([t, t], [L, R], [R, L], [k % 2 == 0, True])
print("group", [len(o["id"]) for o in out])
g.close()
print("sanitize run done")
