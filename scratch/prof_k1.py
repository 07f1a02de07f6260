"""Drives only the event stage (esvio_fe_stage_update: binning + k_sae_update_ts + pyramids) for a
few windows so that ncu can capture the SAE/time-surface kernel in isolation.
usage: python scratch/prof_k1.py [workload] [n_windows]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esvio_b200 import frontend, synth

name = sys.argv[1] if len(sys.argv) > 1 else "stereo_vga_5mevs"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
w = synth.WORKLOADS[name]
cfg = synth.default_config(w["width"], w["height"], max_cnt=w["max_cnt"], min_dist=w["min_dist"])
cfg["max_events_per_window"] = int(w["rate"] / synth.WINDOWS_PER_SEC) + 1024
fe = frontend.EventFrontEnd(cfg)
s = synth.StereoEventStream(w["width"], w["height"], w["rate"], mono=w["mono"])
for k in range(n):
    L, R, t = s.stereo_window(k)
    fe.stage_update(t, L, R)
    fe.stage_corner_flags(L, True)
print("ok", name, n, fe.kernel_launches())
fe.close()
