"""Diagnostics of the frame path on the GPU: per-stage mismatch counts against the oracle and
the cv2 goldens (more detail than the pytest assertions give)."""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esvio_b200 import frontend, synth  # noqa: E402
from oracle import oracle as ora  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "frames_golden.npz"))
imgs = {"tex346": synth.frame_texture(346, 260, 11), "tex640": synth.frame_texture(640, 480, 12),
        "noise173": g["noise173_in"], "flat": np.full((64, 96), 77, np.uint8)}
for name, img in imgs.items():
    try:
        H, W = img.shape
        fe = frontend.EventFrontEnd(dict(synth.default_config(W, H), max_events_per_window=1024))
        mask = g[f"{name}_mask"]
        pts, eig = fe.stage_good_features(img, 100, 30.0, None, want_eig=True)
        ref = ora.corner_min_eigen_val(img)
        bad = np.argwhere(eig != ref)
        print(name, "eig mismatches", len(bad), "cols", sorted(set(bad[:, 1].tolist()))[:10],
              "maxdiff", float(np.abs(eig - ref).max()))
        for tag, (n, md, m) in dict(a=(100, 30.0, None), b=(150, 10.0, mask), c=(0, 1.0, None),
                                    d=(40, 0.5, mask)).items():
            got = fe.stage_good_features(img, n, md, m)
            exp = g[f"{name}_gftt_{tag}"]
            k = min(len(got), len(exp))
            neq = np.nonzero((got[:k] != exp[:k]).any(1))[0]
            print(" ", tag, "n", len(got), "exp", len(exp), "first diff",
                  (int(neq[0]), got[neq[0]].tolist(), exp[neq[0]].tolist()) if len(neq) else None)
        fe.close()
    except Exception:
        traceback.print_exc()

try:
    s = dict(W=240, H=180, n_frames=6, max_cnt=60, min_dist=14)
    cam = [dict(fx=260.0, fy=261.0, cx=121.5, cy=88.0, k1=-0.05, k2=0.02, p1=1e-3, p2=-5e-4),
           dict(fx=259.0, fy=260.5, cx=119.0, cy=90.5, k1=-0.04, k2=0.015, p1=-8e-4, p2=3e-4)]
    cfg = synth.default_config(s["W"], s["H"], max_cnt=s["max_cnt"], min_dist=s["min_dist"])
    cfg["cam"] = cam
    fe = frontend.EventFrontEnd(dict(cfg, max_events_per_window=1024))
    trk = ora.OracleTracker(cfg)
    for k, (L, R) in enumerate(synth.stereo_frame_sequence(s["W"], s["H"], s["n_frames"])):
        right = R if k != 3 else None
        a = fe.track_image(1.0 + k / 20.0, L, right, k % 2 == 0)
        o = trk.track_image(1.0 + k / 20.0, L, right, k % 2 == 0)
        same_ids = np.array_equal(a["id"], o["id"])
        du = float(np.abs(a["u"] - o["u"]).max()) if same_ids and len(a["u"]) else None
        print("frame", k, "n", len(a["id"]), len(o["id"]), "ids equal", same_ids, "max|du|", du,
              "right", len(a["id_right"]), len(o["id_right"]), "stats", a["stats"], o["stats"])
    fe.close()
except Exception:
    traceback.print_exc()
print("frames diag done")
