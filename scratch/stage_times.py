"""Per-stage CUDA-event times of the synchronous call (one window at a time, nothing else on the
GPU): python scratch/stage_times.py [workload] [n_windows]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esvio_b200 import frontend, synth

name = sys.argv[1] if len(sys.argv) > 1 else "stereo_vga_5mevs"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
w = synth.WORKLOADS[name]
cfg = synth.default_config(w["width"], w["height"], max_cnt=w["max_cnt"], min_dist=w["min_dist"])
cfg["max_events_per_window"] = int(w["rate"] / synth.WINDOWS_PER_SEC) + 1024
cfg["use_ransac"] = int(os.environ.get("USE_RANSAC", "1"))
pub_div = int(round(synth.WINDOWS_PER_SEC / w["freq"]))
fe = frontend.EventFrontEnd(cfg)
s = synth.StereoEventStream(w["width"], w["height"], w["rate"], mono=w["mono"])
wins = [s.stereo_window(k) for k in range(n)]
dw = [(frontend._Ev(frontend.DeviceEvents(fe, L)), frontend._Ev(frontend.DeviceEvents(fe, R)), t) for L, R, t in wins]
fe.set_profiling(True)
acc = []
for k, (l, r, t) in enumerate(dw):
    fe.track_raw(t, l, r, k % pub_div == 0)
    if k >= 8:
        acc.append(list(fe.stage_ms().values()))
a = np.array(acc) * 1e3
names = list(fe.stage_ms().keys())
print(name, "K1_DBG=" + os.environ.get("ESVIO_K1_DBG", "0"), " ".join(f"{k}={v:.1f}" for k, v in zip(names, a.mean(0))), "(us, mean)",
      " min: " + " ".join(f"{v:.1f}" for v in a.min(0)))
fe.close()
