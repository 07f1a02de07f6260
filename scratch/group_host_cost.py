"""Host-side cost of a group window: perf_counter around EventFrontEndGroup.submit() / wait() in the
pipelined loop next to the wall time of the loop.  python scratch/group_host_cost.py [workload] [S]"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from esvio_b200 import frontend, synth
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "stereo_vga_5mevs"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 8
K, Wm = 30, 6
w, cfg, pub_div = bench.workload_cfg(wl)
n_per_cam = int(round(w["rate"] / 30))
cfg = dict(cfg, device_id=0, max_events_per_window=n_per_cam + 64)
grp = frontend.EventFrontEndGroup(cfg, S)
m0 = grp.member(0)
bw, held = [], []
for i in range(S):
    row = []
    for L, R, t in bench.gen_windows(w, 100 + i, K + Wm):
        a, b = frontend.DeviceEvents(m0, L), frontend.DeviceEvents(m0, R)
        held += [a, b]
        row.append((frontend._Ev(a), frontend._Ev(b), t))
    bw.append(row)
def gsub(k):
    grp.submit([bw[i][k][2] for i in range(S)], [bw[i][k][0] for i in range(S)],
               [bw[i][k][1] for i in range(S)], [k % pub_div == 0] * S)
for k in range(Wm):
    gsub(k); grp.wait(unpack=False)
torch.cuda.synchronize()
depth = frontend.pipeline_depth()
ts, tw = [], []
t00 = time.perf_counter()
waited = 0
for k in range(Wm, Wm + K):
    t0 = time.perf_counter(); gsub(k); t1 = time.perf_counter()
    ts.append((t1 - t0, k % pub_div == 0))
    if k - Wm >= depth - 1:
        t0 = time.perf_counter(); grp.wait(unpack=False); tw.append(time.perf_counter() - t0); waited += 1
while waited < K:
    t0 = time.perf_counter(); grp.wait(unpack=False); tw.append(time.perf_counter() - t0); waited += 1
torch.cuda.synchronize()
wall = (time.perf_counter() - t00) / K
sp = [a for a, p in ts if p]; sn = [a for a, p in ts if not p]
print(f"{wl} S={S}: wall {wall*1e6:.0f} us/step; submit pub {np.mean(sp)*1e6:.0f} us, non-pub {np.mean(sn)*1e6:.0f} us; "
      f"wait {np.mean(tw)*1e6:.0f} us (min {np.min(tw)*1e6:.0f}); launches/step {grp.kernel_launches()/(K+Wm):.0f}")
grp.close()
