"""Where the synchronous esvio_fe_track call spends its time: per-stage CUDA-event marks of
synchronous windows (profiling mode), relative to the window's submit mark, next to the host
wall time of the call."""
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
if os.environ.get("SYNC_TL_HOST"):   # pinned host buffers, both cameras of a window in one block
    blocks = [frontend.PinnedStereoEvents(L, R) for L, R, _ in wins]
    dw = [(frontend._Ev(b.left), frontend._Ev(b.right), w[2]) for b, w in zip(blocks, wins)]
else:
    dw = [(frontend._Ev(frontend.DeviceEvents(fe, L)), frontend._Ev(frontend.DeviceEvents(fe, R)), t) for L, R, t in wins]
for k in range(Wm):
    fe.submit(dw[k][2], dw[k][0], dw[k][1], k % pub_div == 0); fe.wait(unpack=False)
torch.cuda.synchronize()
for prof in (False, True):
    fe.set_profiling(prof)
    marks, walls, subs = [], [], []
    for k in range(Wm, Wm + K):
        t0 = time.perf_counter()
        fe.submit(dw[k][2], dw[k][0], dw[k][1], k % pub_div == 0)
        t1 = time.perf_counter()
        fe.wait(unpack=False)
        t2 = time.perf_counter()
        walls.append((t2 - t0) * 1e6); subs.append((t1 - t0) * 1e6)
        if prof: marks.append(fe.stage_marks())
    pubs = np.array([k % pub_div == 0 for k in range(Wm, Wm + K)])
    walls, subs = np.array(walls), np.array(subs)
    print(f"{wl} profiling={prof}: wall us/call pub {walls[pubs].mean():.0f} non-pub {walls[~pubs].mean():.0f} | host submit us pub {subs[pubs].mean():.0f} non-pub {subs[~pubs].mean():.0f}")
m = np.array(marks) * 1e3
names = ["submit", "landed", "K1.start", "K1.done", "pyr", "flags", "lk+flt", "select", "packed", "d2h", "T.start", "S.start", "binned"]
order = [0, 1, 12, 2, 3, 4, 5, 10, 6, 7, 11, 8, 9]
rel = m - m[:, :1]
for sel, nm in ((pubs, "publish windows"), (~pubs, "other windows")):
    r = rel[sel].mean(axis=0)
    print(nm + ": " + "  ".join(f"{names[i]} {r[i]:.0f}" for i in order))
