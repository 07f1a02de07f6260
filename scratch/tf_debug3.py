import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esvio_b200 import frontend, synth
from oracle import oracle as ora
W, H, rate, n_windows, pub_every = 346, 260, 1e6, 60, 2
cfg = synth.default_config(W, H, use_ransac=1, max_events_per_window=int(rate / 30) + 64)
fe = frontend.EventFrontEnd(cfg)
ot = ora.OracleTracker(cfg, use_cv2=True, cv2_threads=8)
s = synth.StereoEventStream(W, H, rate)
prev = None
for k in range(n_windows):
    L, R, t = s.stereo_window(k)
    pub = k % pub_every == 0
    if prev is not None:
        fe.stage_set_tracks(prev_time, next_id, prev)
    g = fe.track(t, L, R, pub)
    o = ot.track(t, L, R, pub)
    prev, prev_time, next_id = o, t, ot.next_id()
    ts_diff = [int((fe.time_surface(c) != ot.time_surface(c)).sum()) for c in (0, 1)]
    sae_diff = [sum(int((a != b).sum()) for a, b in zip(fe.sae_planes(c), ot.sae(c).planes())) for c in (0, 1)]
    new_g, new_o = g["track_cnt"] == 1, o["track_cnt"] == 1
    _, ia, ib = np.intersect1d(g["id"][~new_g], o["id"][~new_o], return_indices=True)
    d = np.maximum(np.abs(g["u"][~new_g][ia] - o["u"][~new_o][ib]), np.abs(g["v"][~new_g][ia] - o["v"][~new_o][ib]))
    if k >= 18:
        print(f"k={k} pub={int(pub)} ts diff px {ts_diff} sae diff {sae_diff} old tracks {len(ia)} |d| max {d.max() if len(d) else 0:.2e} >1e-3: {int((d>1e-3).sum())}")
