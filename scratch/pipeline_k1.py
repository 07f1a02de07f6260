"""k_sae_update_ts inside the 3-deep pipeline vs the size of the tracking stages (MAX_CNT):
python scratch/pipeline_k1.py [workload]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esvio_b200 import frontend, synth
name = sys.argv[1] if len(sys.argv) > 1 else "stereo_davis346_1mevs"
w = synth.WORKLOADS[name]
pub_div = int(round(synth.WINDOWS_PER_SEC / w["freq"]))
s = synth.StereoEventStream(w["width"], w["height"], w["rate"])
wins = [s.stereo_window(k) for k in range(70)]
for max_cnt, depth in ((150, 3), (4, 3), (150, 1), (150, 2)):
    cfg = synth.default_config(w["width"], w["height"], max_cnt=max_cnt, min_dist=w["min_dist"], use_ransac=1)
    cfg["max_events_per_window"] = int(w["rate"] / 30) + 1024
    fe = frontend.EventFrontEnd(cfg)
    dw = [(frontend._Ev(frontend.DeviceEvents(fe, L)), frontend._Ev(frontend.DeviceEvents(fe, R)), t) for L, R, t in wins]
    fe.set_profiling(True)
    acc = []
    inflight = 0
    for k, (l, r, t) in enumerate(dw):
        fe.submit(t, l, r, k % pub_div == 0)
        inflight += 1
        if inflight == depth:
            fe.wait(unpack=False); inflight -= 1
            if k > 10:
                acc.append(list(fe.stage_ms().values()))
    while inflight:
        fe.wait(unpack=False); inflight -= 1
    a = np.array(acc) * 1e3
    names = list(fe.stage_ms().keys())
    print(f"{name} max_cnt={max_cnt} depth={depth} " + " ".join(f"{k}={v:.1f}" for k, v in zip(names, a.mean(0))))
    fe.close()
