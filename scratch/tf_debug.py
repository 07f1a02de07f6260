import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esvio_b200 import frontend, synth
from oracle import oracle as ora
W, H, rate, n_windows, pub_every = 346, 260, 1e6, 40, 2
cfg = synth.default_config(W, H, use_ransac=1, max_events_per_window=int(rate / 30) + 64)
fe = frontend.EventFrontEnd(cfg)
ot = ora.OracleTracker(cfg, use_cv2=True, cv2_threads=8)
oc = ora.OracleTracker(cfg, use_cv2=False)   # C-port LK, teacher-forced too? (no set_state: free-running, informational)
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
    if np.array_equal(g["id"], o["id"]) and len(np.setxor1d(g["id_right"], o["id_right"])) > 5:
        only_g = np.setdiff1d(g["id_right"], o["id_right"]); only_o = np.setdiff1d(o["id_right"], g["id_right"])
        print(f"window {k} pub {pub}: left {len(g['id'])} identical; right gpu {len(g['id_right'])} ref {len(o['id_right'])}; only gpu {len(only_g)} only ref {len(only_o)}")
        cnt = dict(zip(o["id"].tolist(), o["track_cnt"].tolist()))
        pos = dict(zip(o["id"].tolist(), zip(o["u"].tolist(), o["v"].tolist())))
        print("  only-gpu ids (track_cnt, left pos):", [(int(i), cnt[int(i)], tuple(round(c,1) for c in pos[int(i)])) for i in only_g[:12]])
        print("  only-ref ids (track_cnt, left pos):", [(int(i), cnt[int(i)], tuple(round(c,1) for c in pos[int(i)])) for i in only_o[:12]])
        # direct LK comparison on this window's images
        a, b = fe.time_surface(0), fe.time_surface(1)
        import cv2
        p0 = np.stack([o["u"], o["v"]], 1).astype(np.float32)
        print('  left |d(u,v)| max', float(np.abs(np.stack([g['u'],g['v']],1)-p0).max()), 'right ids order equal prefix:', int((g['id_right'][:min(len(g['id_right']),len(o['id_right']))]==o['id_right'][:min(len(g['id_right']),len(o['id_right']))]).sum()))
        print('  gpu right ids', g['id_right'][:40].tolist()); print('  ref right ids', o['id_right'][:40].tolist())
        f_cv, st_cv, _ = cv2.calcOpticalFlowPyrLK(a, b, p0.reshape(-1,1,2), None, winSize=(21,21), maxLevel=3)
        f_g, st_g = fe.stage_lk(a, b, p0, None, 3)
        f_cv = f_cv.reshape(-1,2); st_cv = st_cv.reshape(-1)
        both = (st_cv != 0) & (st_g != 0)
        d = np.abs(f_cv - f_g).max(axis=1)
        print(f"  stereo fwd LK on the same images: status differ {int(((st_cv!=0)!=(st_g!=0)).sum())}, |d| max {d[both].max():.3e}, >1e-3: {int((d[both]>1e-3).sum())}")
        r_cv, sr_cv, _ = cv2.calcOpticalFlowPyrLK(b, a, f_cv.reshape(-1,1,2), None, winSize=(21,21), maxLevel=3)
        r_g, sr_g = fe.stage_lk(b, a, f_cv, None, 3)
        r_cv = r_cv.reshape(-1,2); sr_cv = sr_cv.reshape(-1)
        d2 = np.abs(r_cv - r_g).max(axis=1); both2 = (sr_cv!=0)&(sr_g!=0)
        print(f"  stereo bwd LK from cv2's fwd: status differ {int(((sr_cv!=0)!=(sr_g!=0)).sum())}, |d| max {d2[both2].max():.3e}, >1e-3: {int((d2[both2]>1e-3).sum())}")
        fb_cv = np.hypot(*(r_cv - p0).T); fb_g = np.hypot(*(r_g - p0).T)
        print("  fb-dist flips (<=0.5):", int(((fb_cv <= 0.5) != (fb_g <= 0.5)).sum()))
        break
