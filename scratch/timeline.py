"""Gantt chart of the pipeline from the per-stage CUDA events (profiling mode)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from esvio_b200 import frontend
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "stereo_vga_5mevs"
K, Wm = 24, 6
w, cfg, pub_div = bench.workload_cfg(wl)
n_per_cam = int(round(w["rate"] / 30))
cfg = dict(cfg, device_id=0, max_events_per_window=n_per_cam + 64)
wins = bench.gen_windows(w, 0, K + Wm)
fe = frontend.EventFrontEnd(cfg)
dw = [(frontend._Ev(frontend.DeviceEvents(fe, L)), frontend._Ev(frontend.DeviceEvents(fe, R)), t) for L, R, t in wins]
for k in range(Wm):
    fe.submit(dw[k][2], dw[k][0], dw[k][1], k % pub_div == 0); fe.wait(unpack=False)
torch.cuda.synchronize()
fe.set_profiling(True)
marks = []
t0 = time.perf_counter()
for k in range(Wm, Wm + K):
    fe.submit(dw[k][2], dw[k][0], dw[k][1], k % pub_div == 0)
    if k - Wm >= frontend.pipeline_depth() - 1:
        fe.wait(unpack=False); marks.append(fe.stage_marks())
while len(marks) < K:
    fe.wait(unpack=False); marks.append(fe.stage_marks())
torch.cuda.synchronize()
print(f"{wl}: profiled pass wall ms/step {(time.perf_counter()-t0)*1e3/K:.4f}")
m = np.array(marks) * 1e3   # us
print("win pub | submit landed binned | K1.start K1.done | pyr    flags | T.start lk+flt select | S.start packed d2h   (us since profiling on)")
for i, r in enumerate(m):
    k = Wm + i
    print(f"{k:3d} {int(k % pub_div == 0)}   | {r[0]:7.0f} {r[1]:6.0f} {r[12]:6.0f} | {r[2]:7.0f} {r[3]:7.0f} | {r[4]:6.0f} {r[5]:6.0f} | {r[10]:7.0f} {r[6]:6.0f} {r[7]:6.0f} | {r[11]:7.0f} {r[8]:6.0f} {r[9]:6.0f}")
d = np.diff(m[:, 9])
print("period (d2h-done to d2h-done) us: mean %.1f  min %.1f max %.1f" % (d[4:].mean(), d[4:].min(), d[4:].max()))
