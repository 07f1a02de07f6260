"""How far apart are two CPU implementations of the same calcOpticalFlowPyrLK (cv2 4.13 with its
SIMD float sums vs the oracle's scalar C port) on the LK calls of a real run?  That spread is the
floor for any GPU-vs-reference bar on (u, v)."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
from esvio_b200 import synth
from oracle import oracle as ora

def run(W, H, rate, n_windows, pub_every):
    cfg = synth.default_config(W, H, use_ransac=1)
    ot = ora.OracleTracker(cfg, use_cv2=True, cv2_threads=8)
    diffs, flips, total = [], 0, 0
    lk_cv2 = ot._hooks[0]
    def lk(prev, nxt, w, h, pp, npp, n, st, max_level, init):
        nonlocal flips, total
        a = np.ctypeslib.as_array(C.cast(prev, C.POINTER(C.c_uint8)), shape=(H, W)).copy()
        b = np.ctypeslib.as_array(C.cast(nxt, C.POINTER(C.c_uint8)), shape=(H, W)).copy()
        p0 = np.ctypeslib.as_array(pp, shape=(n, 2)).copy()
        p1 = np.ctypeslib.as_array(npp, shape=(n, 2)).copy()
        o_pts, o_st = ora.calc_optical_flow_pyr_lk(a, b, p0, p1 if init else None, max_level=max_level)
        lk_cv2(prev, nxt, w, h, pp, npp, n, st, max_level, init)
        c_pts = np.ctypeslib.as_array(npp, shape=(n, 2)); c_st = np.ctypeslib.as_array(st, shape=(n,))
        both = (c_st != 0) & (o_st != 0)
        flips += int(((c_st != 0) != (o_st != 0)).sum()); total += n
        if both.any(): diffs.append(np.abs(c_pts[both] - o_pts[both]).max(axis=1))
    hook = ora.LK_FN(lk)
    ora.lib().ora_tracker_set_hooks(ot._h, hook, ot._hooks[1], ot._hooks[2])
    s = synth.StereoEventStream(W, H, rate)
    for k in range(n_windows):
        L, R, t = s.stereo_window(k)
        ot.track(t, L, R, k % pub_every == 0)
    d = np.concatenate(diffs)
    print(f"{W}x{H}: {total} LK point-calls, status flips {flips}; |cv2 - C port| median {np.median(d):.2e} 99% {np.quantile(d,0.99):.2e} 99.9% {np.quantile(d,0.999):.2e} max {d.max():.2e}; >1e-3: {(d>1e-3).sum()}  >1e-2: {(d>1e-2).sum()} >0.1: {(d>0.1).sum()}")
run(346, 260, 1e6, 90, 2)
run(640, 480, 5e6, 45, 3)
