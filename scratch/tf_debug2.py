import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
from esvio_b200 import frontend, synth
from oracle import oracle as ora
W, H, rate, n_windows, pub_every = 346, 260, 1e6, 70, 2
cfg = synth.default_config(W, H, use_ransac=1, max_events_per_window=int(rate / 30) + 64)
fe = frontend.EventFrontEnd(cfg)
ot = ora.OracleTracker(cfg, use_cv2=True, cv2_threads=8)
s = synth.StereoEventStream(W, H, rate)
prev_img = None
tot = {"gc": [], "go": [], "co": []}
for k in range(n_windows):
    L, R, t = s.stereo_window(k)
    pub = k % pub_every == 0
    p_prev = None if k == 0 else np.stack([o["u"], o["v"]], 1).astype(np.float32)
    o = ot.track(t, L, R, pub)
    cur_img = ot.lk_image(0)
    if k >= 40 and p_prev is not None and len(p_prev):
        f_cv, st_cv, _ = cv2.calcOpticalFlowPyrLK(prev_img, cur_img, p_prev.reshape(-1,1,2), None, winSize=(21,21), maxLevel=3)
        f_cv = f_cv.reshape(-1,2); st_cv = st_cv.reshape(-1) != 0
        f_g, st_g = fe.stage_lk(prev_img, cur_img, p_prev, None, 3); st_g = st_g != 0
        f_o, st_o = ora.calc_optical_flow_pyr_lk(prev_img, cur_img, p_prev, None, max_level=3); st_o = st_o != 0
        for key, (a, sa, b, sb) in {"gc": (f_g, st_g, f_cv, st_cv), "go": (f_g, st_g, f_o, st_o), "co": (f_cv, st_cv, f_o, st_o)}.items():
            both = sa & sb
            tot[key].append(np.abs(a[both] - b[both]).max(axis=1))
        if k in (45, 60):
            d = np.abs(f_g - f_cv).max(axis=1)
            idx = np.argsort(-d)[:6]
            for i in idx:
                print(f"  k={k} pt {p_prev[i]} gpu {f_g[i]} cv2 {f_cv[i]} cport {f_o[i]} |g-cv| {d[i]:.2e} flow {np.hypot(*(f_cv[i]-p_prev[i])):.2f}")
    prev_img = cur_img
for key, name in (("gc", "GPU vs cv2"), ("go", "GPU vs C port"), ("co", "cv2 vs C port")):
    d = np.concatenate(tot[key])
    print(f"{name}: n {len(d)} median {np.median(d):.1e} 99% {np.quantile(d,0.99):.1e} 99.9% {np.quantile(d,0.999):.1e} max {d.max():.1e} >1e-3 {int((d>1e-3).sum())} >1e-2 {int((d>1e-2).sum())}")
