"""Host-side cost of one window: perf_counter around submit() and wait() in the pipelined loop,
next to the CUDA-event time of the same loop."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from esvio_b200 import frontend, synth
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "stereo_vga_5mevs"
K, Wm = 60, 6
w, cfg, pub_div = bench.workload_cfg(wl)
n_per_cam = int(round(w["rate"] / 30))
cfg = dict(cfg, device_id=0, max_events_per_window=n_per_cam + 64)
wins = bench.gen_windows(w, 0, K + Wm)
fe = frontend.EventFrontEnd(cfg)
dw = [(frontend._Ev(frontend.DeviceEvents(fe, L)), frontend._Ev(frontend.DeviceEvents(fe, R)), t) for L, R, t in wins]
for k in range(Wm):
    fe.submit(dw[k][2], dw[k][0], dw[k][1], k % pub_div == 0); fe.wait(unpack=False)
torch.cuda.synchronize()
ext = torch.cuda.ExternalStream(fe.stream())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts, tw = [], []
e0.record(ext)
t00 = time.perf_counter()
for k in range(Wm, Wm + K):
    t0 = time.perf_counter()
    fe.submit(dw[k][2], dw[k][0], dw[k][1], k % pub_div == 0)
    t1 = time.perf_counter()
    ts.append((t1 - t0, k % pub_div == 0))
    if k - Wm >= frontend.pipeline_depth() - 1:
        fe.wait(unpack=False)
        tw.append(time.perf_counter() - t1)
while len(tw) < K:
    t1 = time.perf_counter(); fe.wait(unpack=False); tw.append(time.perf_counter() - t1)
e1.record(ext)
torch.cuda.synchronize()
wall = time.perf_counter() - t00
pub = [a for a, p in ts if p]; non = [a for a, p in ts if not p]
print(f"{wl}: gpu ms/step {e0.elapsed_time(e1)/K:.4f} wall ms/step {wall*1e3/K:.4f} | submit us: pub {np.mean(pub)*1e6:.1f} non-pub {np.mean(non)*1e6:.1f} | wait us mean {np.mean(tw)*1e6:.1f} median {np.median(tw)*1e6:.1f}")
# submit-only cost: no waits in between beyond the pipeline depth (host never blocks on the GPU if GPU is faster)
