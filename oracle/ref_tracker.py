"""The reference's own FeatureTracker behind ctypes (TEST INFRASTRUCTURE, like everything under
oracle/: only tests/ and bench.py's CPU arm may use it).

oracle/_ref/libesvio_ref_ft.so = /root/reference's feature_tracker.cpp + event_detector.cc
compiled unmodified (oracle/Makefile, oracle/ref_shim/ref_ft_api.cc).  Used by
tests/test_oracle_ref_tracker.py (oracle == reference code, CPU), tests/test_gpu_parity.py
(CUDA == reference code, on the GPU box: the prebuilt library travels with the snapshot, the
reference tree does not) and `bench.py --impl reference` (an informational timing)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libesvio_ref_ft.so")
REF_SRC = "/root/reference/feature_tracker/src/feature_tracker.cpp"

_p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
KEYS_L = ("id", "track_cnt", "u", "v", "un_x", "un_y", "vx", "vy")
KEYS_R = ("id_right", "ru", "rv", "run_x", "run_y", "rvx", "rvy")


def load():
    """The library, (re)built first where the reference tree is present; None where neither the
    tree nor a prebuilt library exists."""
    if os.path.exists(REF_SRC):  # this container: (re)build from the reference where it lies
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref"] + os.environ.get("ESVIO_REF_MAKE_ARGS", "").split())
    if not os.path.exists(REF_SO):
        return None
    L = C.CDLL(REF_SO)
    L.ref_ft_create.restype = C.c_void_p
    L.ref_ft_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.ref_ft_destroy.argtypes = [C.c_void_p]
    L.ref_ft_track.argtypes = ([C.c_void_p, C.c_double] + [C.c_void_p] * 5 + [C.c_size_t]
                               + [C.c_void_p] * 5 + [C.c_size_t, C.c_int, C.c_uint32, C.c_uint32]
                               + [C.c_void_p] * 4)
    L.ref_ft_track_image.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int]
    L.ref_ft_counts.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_ft_get.argtypes = [C.c_void_p] + [C.c_void_p] * 9
    L.ref_ft_lk_image.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.ref_ft_set_intrinsics.argtypes = [C.c_double] * 4
    L.ref_std_sort_order.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return L


_LK_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                       C.c_int, C.POINTER(C.c_uint8), C.c_int, C.c_int)
_FM_HOOK = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_double, C.POINTER(C.c_uint8))
_IMG_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_void_p)
_GFTT_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double,
                         C.POINTER(C.c_float))
_installed = {}   # library handle -> the callback objects (ctypes callbacks must outlive their use)


def use_real_opencv(L, on=True, threads=None):
    """Serve the reference code's OpenCV calls (calcOpticalFlowPyrLK, findFundamentalMat, CLAHE,
    normalize, goodFeaturesToTrack -- mini_cv.h) with REAL OpenCV through cv2 instead of the
    oracle's restatements, with the arguments feature_tracker.cpp passes (:410,417-418,490,495,
    935,228,377-381).  `on=False` restores the restatements."""
    L.ref_ft_set_cv_hooks.argtypes = [C.c_void_p] * 5
    if not on:
        L.ref_ft_set_cv_hooks(None, None, None, None, None)
        _installed.pop(id(L), None)
        return
    import cv2
    if threads is not None:
        cv2.setNumThreads(threads)
    u8 = C.POINTER(C.c_uint8)

    def img(p, w, h):
        return np.ctypeslib.as_array(C.cast(p, u8), shape=(h, w))

    def lk(prev, nxt, w, h, pp, npp, n, st, max_level, init):
        a, b = img(prev, w, h), img(nxt, w, h)
        p0 = np.ctypeslib.as_array(pp, shape=(n, 2))
        p1 = np.ctypeslib.as_array(npp, shape=(n, 2))
        s = np.ctypeslib.as_array(st, shape=(n,))
        if init:
            out, status, _ = cv2.calcOpticalFlowPyrLK(
                a, b, p0.reshape(-1, 1, 2), p1.reshape(-1, 1, 2).copy(), winSize=(21, 21), maxLevel=max_level,
                criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        else:
            out, status, _ = cv2.calcOpticalFlowPyrLK(a, b, p0.reshape(-1, 1, 2), None, winSize=(21, 21),
                                                      maxLevel=max_level)
        p1[:] = out.reshape(-1, 2)
        s[:] = status.reshape(-1)

    def fm(a, b, n, thr, mask):
        m = np.ctypeslib.as_array(mask, shape=(n,))
        _F, status = cv2.findFundamentalMat(np.ctypeslib.as_array(a, shape=(n, 2)),
                                            np.ctypeslib.as_array(b, shape=(n, 2)), cv2.FM_RANSAC, thr, 0.99)
        if status is None:
            m[:] = 0
            return 0
        m[:] = status.reshape(-1)
        return 1

    def clahe(src, w, h, dst):
        img(dst, w, h)[:] = cv2.createCLAHE().apply(img(src, w, h))

    def normalize(src, w, h, dst):
        img(dst, w, h)[:] = cv2.normalize(img(src, w, h), None, 0, 255, cv2.NORM_MINMAX)

    def gftt(im, w, h, mask, max_corners, quality, min_dist, out_xy):
        m = img(mask, w, h) if mask else None
        c = cv2.goodFeaturesToTrack(img(im, w, h), max_corners, quality, min_dist, mask=m)
        if c is None:
            return 0
        c = c.reshape(-1, 2).astype(np.float32)
        np.ctypeslib.as_array(out_xy, shape=(len(c), 2))[:] = c
        return len(c)

    cbs = (_LK_HOOK(lk), _FM_HOOK(fm), _IMG_HOOK(clahe), _IMG_HOOK(normalize), _GFTT_HOOK(gftt))
    _installed[id(L)] = cbs
    L.ref_ft_set_cv_hooks(*[C.cast(cb, C.c_void_p) for cb in cbs])


class RefTracker:
    """The reference's FeatureTracker behind oracle/ref_shim/ref_ft_api.cc (one at a time:
    `detector`, n_id and the parameters are process-wide globals in the reference)."""

    def __init__(self, L, cfg):
        self.L, self.cfg = L, cfg
        icfg = np.array([cfg["width"], cfg["height"], cfg["max_cnt"], cfg["min_dist"], cfg["flow_back"],
                         cfg["equalize"], cfg["ignore_polarity"], cfg["median_blur_kernel_size"],
                         int(cfg["focal_length"])], np.int32)
        d = [cfg["f_threshold"], cfg["ts_lk_threshold"], cfg["decay_ms"], cfg["feature_filter_threshold"]]
        for cam in cfg["cam"]:
            d += [cam[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")]
        dcfg = np.array(d, np.float64)
        self.h = L.ref_ft_create(_p(icfg), _p(dcfg), 0)

    def close(self):
        if self.h:
            self.L.ref_ft_destroy(self.h)
            self.h = None

    def _results(self):
        c = np.zeros(10, np.int32)
        self.L.ref_ft_counts(self.h, _p(c))
        nl, nr = int(c[0]), int(c[1])
        # the reference's vectors are index-aligned by construction
        assert c[3] == c[4] == c[5] == c[6] == nl, c
        # ... except right_pts_velocity while prev_un_right_pts_map is empty: ptsVelocity then pushes
        # cur_pts.size() zeros -- the LEFT count (feature_tracker.cpp:1037-1043); the node reads the
        # first ids_right.size() of them (stereo_event_tracker_node.cpp:316-323)
        assert c[7] == c[8] == nr and c[9] in (nr, nl), c
        ids, cnt = np.zeros(nl, np.int32), np.zeros(nl, np.int32)
        pts, un, vel = (np.zeros((nl, 2), np.float32) for _ in range(3))
        idr = np.zeros(nr, np.int32)
        rp, run = (np.zeros((nr, 2), np.float32) for _ in range(2))
        rv = np.zeros((int(c[9]), 2), np.float32)
        self.L.ref_ft_get(self.h, _p(ids), _p(cnt), _p(pts), _p(un), _p(vel), _p(idr), _p(rp), _p(run), _p(rv))
        return {"id": ids, "track_cnt": cnt, "u": pts[:, 0], "v": pts[:, 1], "un_x": un[:, 0], "un_y": un[:, 1],
                "vx": vel[:, 0], "vy": vel[:, 1], "id_right": idr, "ru": rp[:, 0], "rv": rp[:, 1],
                "run_x": run[:, 0], "run_y": run[:, 1], "rvx": rv[:nr, 0], "rvy": rv[:nr, 1], "next_id": int(c[2])}

    def track(self, cur_time, L6, R6, pub, stamp=None, motion=None):
        lx, ly, _, lp, lsec, lnsec = (np.ascontiguousarray(a) for a in L6)
        rx, ry, _, rp, rsec, rnsec = (np.ascontiguousarray(a) for a in R6)
        st = stamp if stamp is not None else (0, 0)
        margs = [None] * 4
        keep = []
        if motion is not None:
            keep = [np.array(list(motion["state_v"]) + [0.0], np.float64), np.array(motion["v_pre"], np.float32),
                    np.array(motion["accel"], np.float32), np.array(motion["omega"], np.float32)]
            margs = [_p(a) for a in keep]
        self.L.ref_ft_track(self.h, float(cur_time), _p(lx), _p(ly), _p(lsec), _p(lnsec), _p(lp), len(lx),
                            _p(rx), _p(ry), _p(rsec), _p(rnsec), _p(rp), len(rx), int(pub), st[0], st[1], *margs)
        return self._results()

    def track_image(self, cur_time, left, right, pub):
        left = np.ascontiguousarray(left)
        right = None if right is None else np.ascontiguousarray(right)
        self.L.ref_ft_track_image(self.h, float(cur_time), _p(left), None if right is None else _p(right), int(pub))
        return self._results()

    def lk_image(self, cam):
        out = np.zeros((self.cfg["height"], self.cfg["width"]), np.uint8)
        self.L.ref_ft_lk_image(self.h, cam, _p(out))
        return out


