"""ctypes wrapper around oracle/libesvio_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module (see oracle/esvio_oracle.h).  It exposes
the stage functions of the CPU restatement plus `OracleTracker`, the per-window
orchestration of FeatureTracker::trackEvent
(/root/reference/feature_tracker/src/feature_tracker.cpp:340-603).

`OracleTracker(use_cv2=True)` routes the OpenCV stages (calcOpticalFlowPyrLK,
findFundamentalMat, CLAHE) through real OpenCV via cv2 when it is importable;
`use_cv2=False` uses the C ports in esvio_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libesvio_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "esvio_oracle.c")
    hdr = os.path.join(_HERE, "esvio_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class _Pinhole(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")]


class _Config(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int), ("max_cnt", C.c_int), ("min_dist", C.c_int),
        ("flow_back", C.c_int), ("equalize", C.c_int),
        ("f_threshold", C.c_double), ("ts_lk_threshold", C.c_double), ("decay_ms", C.c_double),
        ("ignore_polarity", C.c_int), ("median_blur_kernel_size", C.c_int),
        ("feature_filter_threshold", C.c_double), ("focal_length", C.c_double),
        ("cam", _Pinhole * 2),
    ]


_pf = C.POINTER(C.c_float)
_pi = C.POINTER(C.c_int)


class _Tracks(C.Structure):
    _fields_ = [
        ("n_left", C.c_int), ("id", _pi), ("track_cnt", _pi),
        ("u", _pf), ("v", _pf), ("un_x", _pf), ("un_y", _pf), ("vx", _pf), ("vy", _pf),
        ("n_right", C.c_int), ("id_right", _pi),
        ("ru", _pf), ("rv", _pf), ("run_x", _pf), ("run_y", _pf), ("rvx", _pf), ("rvy", _pf),
        ("n_prev", C.c_int), ("n_after_temporal", C.c_int), ("n_after_ransac", C.c_int),
        ("n_after_mask", C.c_int), ("n_new", C.c_int),
    ]


class _Motion(C.Structure):
    _fields_ = [("state_v", C.c_double * 3), ("v_pre", C.c_float * 3), ("accel", C.c_float * 3),
                ("omega", C.c_float * 3), ("t1", C.c_double), ("K", C.c_float * 4)]


def make_motion(m: dict) -> "_Motion":
    """dict(state_v, v_pre, accel, omega, t1, K=(fx, fy, cx, cy)) -> ora_motion."""
    o = _Motion()
    for k in ("state_v", "v_pre", "accel", "omega"):
        for i in range(3):
            getattr(o, k)[i] = float(m[k][i])
    o.t1 = float(m["t1"])
    for i in range(4):
        o.K[i] = float(m["K"][i])
    return o


class _Sae(C.Structure):
    _fields_ = [("W", C.c_int), ("H", C.c_int), ("sae", C.POINTER(C.c_double) * 2),
                ("latest", C.POINTER(C.c_double) * 2)]


LK_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_int, _pf, _pf, C.c_int,
                    C.POINTER(C.c_uint8), C.c_int, C.c_int)
FM_FN = C.CFUNCTYPE(C.c_int, _pf, _pf, C.c_int, C.c_double, C.POINTER(C.c_uint8))
EQ_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_void_p)

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ora_sae_create.restype = C.POINTER(_Sae)
        L.ora_sae_create.argtypes = [C.c_int, C.c_int]
        L.ora_sae_destroy.argtypes = [C.POINTER(_Sae)]
        L.ora_sae_reset.argtypes = [C.POINTER(_Sae)]
        L.ora_sae_update.argtypes = [C.POINTER(_Sae), C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.c_double]
        L.ora_time_surface.argtypes = [C.POINTER(_Sae), C.c_double, C.c_double, C.c_int, C.c_void_p]
        L.ora_is_corner.restype = C.c_int
        L.ora_is_corner.argtypes = [C.POINTER(_Sae), C.c_double, C.c_int, C.c_int, C.c_int,
                                    C.c_double, C.c_int]
        L.ora_corner_flags.argtypes = [C.POINTER(_Sae), C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_void_p]
        L.ora_disc_half_widths.argtypes = [C.c_int, C.c_void_p]
        L.ora_fill_disc_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_uint8]
        L.ora_set_mask.restype = C.c_int
        L.ora_std_sort_order.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ora_set_mask.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]
        L.ora_features_to_track.restype = C.c_int
        L.ora_features_to_track.argtypes = [C.POINTER(_Sae), C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_double, C.c_double, C.c_void_p,
                                            C.c_void_p]
        L.ora_pyramid_sizes.restype = C.c_int
        L.ora_pyramid_sizes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ora_pyr_down.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.ora_scharr_deriv.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ora_calc_optical_flow_pyr_lk.argtypes = [
            C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
            C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double]
        L.ora_lift_projective.argtypes = [C.POINTER(_Pinhole), C.c_double, C.c_double,
                                          C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ora_solve_cubic.restype = C.c_int
        L.ora_solve_cubic.argtypes = [C.c_void_p, C.c_void_p]
        L.ora_run_7point.restype = C.c_int
        L.ora_run_7point.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ora_find_fundamental_mask.restype = C.c_int
        L.ora_find_fundamental_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                                C.c_double, C.c_int, C.c_void_p]
        L.ora_median_blur_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ora_clahe_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p]
        L.ora_normalize_minmax_u8.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.ora_equalize_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ora_corner_min_eigen_val_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ora_good_features_to_track.restype = C.c_int
        L.ora_good_features_to_track.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                                 C.c_double, C.c_double, C.c_void_p]
        L.ora_tracker_track_image.restype = C.c_int
        L.ora_tracker_track_image.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                              C.c_int, C.POINTER(_Tracks)]
        L.ora_mat3_exp_f.argtypes = [C.c_void_p, C.c_void_p]
        L.ora_motion_correct.argtypes = [C.POINTER(_Motion), C.c_int, C.c_int, C.c_double, C.c_double,
                                         C.c_double, _pi, _pi]
        L.ora_motion_active.restype = C.c_int
        L.ora_motion_active.argtypes = [C.POINTER(_Motion)]
        L.ora_sae_update_mc.argtypes = [C.POINTER(_Sae), C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_size_t, C.c_double, C.POINTER(_Motion),
                                        C.c_double]
        L.ora_tracker_track_mc.restype = C.c_int
        L.ora_tracker_track_mc.argtypes = [C.c_void_p, C.c_double] + [C.c_void_p] * 4 + [C.c_size_t] + \
            [C.c_void_p] * 4 + [C.c_size_t, C.c_int, C.POINTER(_Motion), C.POINTER(_Tracks)]
        L.ora_tracker_create.restype = C.c_void_p
        L.ora_tracker_create.argtypes = [C.POINTER(_Config)]
        L.ora_tracker_destroy.argtypes = [C.c_void_p]
        L.ora_tracker_set_hooks.argtypes = [C.c_void_p, LK_FN, FM_FN, EQ_FN]
        L.ora_tracker_disable_ransac.argtypes = [C.c_void_p, C.c_int]
        L.ora_tracker_track.restype = C.c_int
        L.ora_tracker_track.argtypes = [C.c_void_p, C.c_double] + [C.c_void_p] * 4 + [C.c_size_t] + \
            [C.c_void_p] * 4 + [C.c_size_t, C.c_int, C.POINTER(_Tracks)]
        L.ora_tracker_sae.restype = C.POINTER(_Sae)
        L.ora_tracker_sae.argtypes = [C.c_void_p, C.c_int]
        L.ora_tracker_time_surface.restype = C.POINTER(C.c_uint8)
        L.ora_tracker_time_surface.argtypes = [C.c_void_p, C.c_int]
        L.ora_tracker_lk_image.restype = C.POINTER(C.c_uint8)
        L.ora_tracker_lk_image.argtypes = [C.c_void_p, C.c_int]
        L.ora_tracker_timers.argtypes = [C.c_void_p, C.c_void_p]
        L.ora_tracker_next_id.restype = C.c_int
        L.ora_tracker_next_id.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _ev(x, y, t, p):
    return (np.ascontiguousarray(x, np.uint16), np.ascontiguousarray(y, np.uint16),
            np.ascontiguousarray(t, np.float64), np.ascontiguousarray(p, np.uint8))


# --------------------------------------------------------------------------- SAE
class Sae:
    """Per-camera SAE state (event_detector.h:74-79) with the stage functions."""

    def __init__(self, W, H, _borrow=None):
        self.W, self.H = W, H
        self._own = _borrow is None
        self._h = lib().ora_sae_create(W, H) if self._own else _borrow

    def __del__(self):
        if getattr(self, "_own", False) and self._h:
            lib().ora_sae_destroy(self._h)
            self._h = None

    def update(self, x, y, t, p, filter_threshold=0.01):
        x, y, t, p = _ev(x, y, t, p)
        lib().ora_sae_update(self._h, _p(x), _p(y), _p(t), _p(p), len(x), filter_threshold)

    def update_mc(self, x, y, t, p, motion: dict, t0, filter_threshold=0.01):
        """createSAE_*(..., measurements) for every event (feature_tracker.cpp:628-642)."""
        x, y, t, p = _ev(x, y, t, p)
        m = make_motion(motion)
        lib().ora_sae_update_mc(self._h, _p(x), _p(y), _p(t), _p(p), len(x), filter_threshold,
                                C.byref(m), float(t0))

    def planes(self):
        """(sae[0], sae[1], latest[0], latest[1]) as HxW float64 copies."""
        s = self._h.contents
        n = self.W * self.H
        out = []
        for arr in (s.sae, s.latest):
            for k in range(2):
                out.append(np.ctypeslib.as_array(arr[k], shape=(n,)).reshape(self.H, self.W).copy())
        return out

    def time_surface(self, t_ref, decay_ms=20.0, ignore_polarity=0):
        out = np.empty((self.H, self.W), np.uint8)
        lib().ora_time_surface(self._h, t_ref, decay_ms, ignore_polarity, _p(out))
        return out

    def is_corner(self, t, x, y, p, filter_threshold=0.01, min_dist=10):
        return bool(lib().ora_is_corner(self._h, t, int(x), int(y), int(p), filter_threshold, min_dist))

    def corner_flags(self, x, y, t, p, filter_threshold=0.01, min_dist=10):
        x, y, t, p = _ev(x, y, t, p)
        out = np.zeros(len(x), np.uint8)
        lib().ora_corner_flags(self._h, _p(x), _p(y), _p(t), _p(p), len(x), filter_threshold,
                               min_dist, _p(out))
        return out

    def features_to_track(self, x, y, t, p, max_corners, min_dist, mask, ts, ts_lk_threshold=128.0,
                          filter_threshold=0.01):
        x, y, t, p = _ev(x, y, t, p)
        mask = np.ascontiguousarray(mask, np.uint8)
        ts = np.ascontiguousarray(ts, np.uint8)
        out = np.zeros((max(max_corners, 1), 2), np.float32)
        mask_out = np.zeros_like(mask)
        k = lib().ora_features_to_track(self._h, _p(x), _p(y), _p(t), _p(p), len(x), max_corners,
                                        min_dist, _p(mask), _p(ts), ts_lk_threshold,
                                        filter_threshold, _p(out), _p(mask_out))
        return out[:k].copy(), mask_out


# --------------------------------------------------------------------------- motion compensation
def mat3_exp_f(A):
    A = np.ascontiguousarray(A, np.float32).reshape(3, 3)
    R = np.empty((3, 3), np.float32)
    lib().ora_mat3_exp_f(_p(A), _p(R))
    return R


def motion_correct(motion: dict, W, H, ex, ey, dt):
    m = make_motion(motion)
    ox, oy = C.c_int(), C.c_int()
    lib().ora_motion_correct(C.byref(m), W, H, float(ex), float(ey), float(dt), C.byref(ox), C.byref(oy))
    return ox.value, oy.value


def motion_active(motion: dict) -> bool:
    m = make_motion(motion)
    return bool(lib().ora_motion_active(C.byref(m)))


# --------------------------------------------------------------------------- image conditioning
def median_blur(img, ksize):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().ora_median_blur_u8(_p(img), img.shape[1], img.shape[0], int(ksize), _p(out))
    return out


def clahe(img, clip_limit=40.0, tiles=8):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().ora_clahe_u8(_p(img), img.shape[1], img.shape[0], float(clip_limit), int(tiles), _p(out))
    return out


def normalize_minmax(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().ora_normalize_minmax_u8(_p(img), img.size, _p(out))
    return out


def corner_min_eigen_val(img):
    """cv2.cornerMinEigenVal(img, 3, ksize=3) on CV_8U."""
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty(img.shape, np.float32)
    lib().ora_corner_min_eigen_val_u8(_p(img), img.shape[1], img.shape[0], _p(out))
    return out


def good_features_to_track(img, max_corners, quality=0.01, min_distance=30.0, mask=None):
    """cv2.goodFeaturesToTrack(img, max_corners, quality, min_distance, mask=mask) -> (n, 2) f32."""
    img = np.ascontiguousarray(img, np.uint8)
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    out = np.zeros((max(int(max_corners), 1) if max_corners > 0 else img.size, 2), np.float32)
    n = lib().ora_good_features_to_track(_p(img), img.shape[1], img.shape[0],
                                         _p(m) if m is not None else None, int(max_corners),
                                         float(quality), float(min_distance), _p(out))
    return out[:n].copy()


def equalize(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().ora_equalize_u8(_p(img), img.shape[1], img.shape[0], _p(out))
    return out


# --------------------------------------------------------------------------- helpers
def disc_half_widths(r):
    out = np.zeros(r + 1, np.int32)
    lib().ora_disc_half_widths(r, _p(out))
    return out


def fill_disc(mask, cx, cy, r, value=255):
    H, W = mask.shape
    lib().ora_fill_disc_u8(_p(mask), W, H, cx, cy, r, value)


def std_sort_order(key, depth_limit=-1):
    """order[k] = index libstdc++'s std::sort (key descending) leaves at position k."""
    key = np.ascontiguousarray(key, np.int32)
    out = np.zeros(len(key), np.int32)
    lib().ora_std_sort_order(_p(key), len(key), int(depth_limit), _p(out))
    return out


def set_mask(W, H, min_dist, pts, ids, track_cnt):
    pts = np.ascontiguousarray(pts, np.float32).copy()
    ids = np.ascontiguousarray(ids, np.int32).copy()
    cnt = np.ascontiguousarray(track_cnt, np.int32).copy()
    mask = np.zeros((H, W), np.uint8)
    m = lib().ora_set_mask(W, H, min_dist, len(ids), _p(pts), _p(ids), _p(cnt), _p(mask))
    return pts[:m], ids[:m], cnt[:m], mask


def pyramid_sizes(W, H, max_level=3, win=21):
    w = np.zeros(16, np.int32)
    h = np.zeros(16, np.int32)
    n = lib().ora_pyramid_sizes(W, H, max_level, win, _p(w), _p(h))
    return [(int(w[i]), int(h[i])) for i in range(n)]


def pyr_down(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().ora_pyr_down(_p(img), w, h, _p(out), out.shape[1], out.shape[0])
    return out


def build_pyramid(img, max_level=3, win=21):
    levels = [np.ascontiguousarray(img, np.uint8)]
    for _ in range(len(pyramid_sizes(img.shape[1], img.shape[0], max_level, win)) - 1):
        levels.append(pyr_down(levels[-1]))
    return levels


def scharr_deriv(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty((h, w, 2), np.int16)
    lib().ora_scharr_deriv(_p(img), w, h, _p(out))
    return out


def calc_optical_flow_pyr_lk(prev, nxt, prev_pts, next_pts=None, max_level=3, win=21,
                             max_count=30, epsilon=0.01, min_eig=1e-4):
    prev = np.ascontiguousarray(prev, np.uint8)
    nxt = np.ascontiguousarray(nxt, np.uint8)
    pp = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
    n = len(pp)
    init = next_pts is not None
    npts = np.ascontiguousarray(next_pts, np.float32).reshape(-1, 2).copy() if init \
        else np.zeros((n, 2), np.float32)
    st = np.zeros(n, np.uint8)
    H, W = prev.shape
    lib().ora_calc_optical_flow_pyr_lk(_p(prev), _p(nxt), W, H, _p(pp), _p(npts), n, _p(st), win,
                                       max_level, max_count, epsilon, int(init), min_eig)
    return npts, st


def lift_projective(cam, u, v):
    c = _Pinhole(*[float(cam[k]) for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")])
    x, y = C.c_double(), C.c_double()
    lib().ora_lift_projective(C.byref(c), float(u), float(v), C.byref(x), C.byref(y))
    return x.value, y.value


def solve_cubic(c):
    c = np.ascontiguousarray(c, np.float64)
    r = np.zeros(3, np.float64)
    n = lib().ora_solve_cubic(_p(c), _p(r))
    return n, r


def run_7point(m1, m2):
    m1 = np.ascontiguousarray(m1, np.float32)
    m2 = np.ascontiguousarray(m2, np.float32)
    F = np.zeros((3, 3, 3), np.float64)
    n = lib().ora_run_7point(_p(m1), _p(m2), _p(F))
    return F[:max(n, 0)]


def find_fundamental_mask(p1, p2, thresh=1.0, confidence=0.99, max_iters=1000):
    p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 2)
    mask = np.zeros(len(p1), np.uint8)
    ok = lib().ora_find_fundamental_mask(_p(p1), _p(p2), len(p1), thresh, confidence, max_iters,
                                         _p(mask))
    return bool(ok), mask


# --------------------------------------------------------------------------- tracker
def have_cv2():
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


def make_config(cfg: dict) -> _Config:
    c = _Config()
    c.width, c.height = cfg["width"], cfg["height"]
    c.max_cnt, c.min_dist = cfg.get("max_cnt", 150), cfg.get("min_dist", 10)
    c.flow_back, c.equalize = cfg.get("flow_back", 1), cfg.get("equalize", 0)
    c.f_threshold = cfg.get("f_threshold", 1.0)
    c.ts_lk_threshold = cfg.get("ts_lk_threshold", 128.0)
    c.decay_ms = cfg.get("decay_ms", 20.0)
    c.ignore_polarity = cfg.get("ignore_polarity", 0)
    c.median_blur_kernel_size = cfg.get("median_blur_kernel_size", 0)
    c.feature_filter_threshold = cfg.get("feature_filter_threshold", 0.01)
    c.focal_length = cfg.get("focal_length", 460.0)
    for i in range(2):
        cam = cfg["cam"][i]
        c.cam[i] = _Pinhole(*[float(cam[k]) for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")])
    return c


class OracleTracker:
    """FeatureTracker::trackEvent (feature_tracker.cpp:340-603) on the CPU."""

    def __init__(self, cfg: dict, use_cv2: bool = False, cv2_threads: int | None = None,
                 disable_ransac: bool = False):
        self.cfg = dict(cfg)
        self._c = make_config(cfg)
        self._h = lib().ora_tracker_create(C.byref(self._c))
        self.W, self.H, self.M = cfg["width"], cfg["height"], cfg.get("max_cnt", 150)
        self.use_cv2 = bool(use_cv2 and have_cv2())
        self._hooks = None
        if disable_ransac:
            lib().ora_tracker_disable_ransac(self._h, 1)
        if self.use_cv2:
            import cv2
            if cv2_threads is not None:
                cv2.setNumThreads(cv2_threads)
            W, H = self.W, self.H

            def lk(prev, nxt, w, h, pp, npp, n, st, max_level, init):
                a = np.ctypeslib.as_array(C.cast(prev, C.POINTER(C.c_uint8)), shape=(H, W))
                b = np.ctypeslib.as_array(C.cast(nxt, C.POINTER(C.c_uint8)), shape=(H, W))
                p0 = np.ctypeslib.as_array(pp, shape=(n, 2))
                p1 = np.ctypeslib.as_array(npp, shape=(n, 2))
                s = np.ctypeslib.as_array(st, shape=(n,))
                if init:
                    out, status, _ = cv2.calcOpticalFlowPyrLK(
                        a, b, p0.reshape(-1, 1, 2), p1.reshape(-1, 1, 2).copy(), winSize=(21, 21),
                        maxLevel=max_level,
                        criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                        flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
                else:
                    out, status, _ = cv2.calcOpticalFlowPyrLK(
                        a, b, p0.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=max_level)
                p1[:] = out.reshape(-1, 2)
                s[:] = status.reshape(-1)

            def fm(a, b, n, thr, mask):
                p0 = np.ctypeslib.as_array(a, shape=(n, 2))
                p1 = np.ctypeslib.as_array(b, shape=(n, 2))
                m = np.ctypeslib.as_array(mask, shape=(n,))
                F, status = cv2.findFundamentalMat(p0, p1, cv2.FM_RANSAC, thr, 0.99)
                if status is None:
                    m[:] = 0
                    return 0
                m[:] = status.reshape(-1)
                return 1

            def eq(src, w, h, dst):
                a = np.ctypeslib.as_array(C.cast(src, C.POINTER(C.c_uint8)), shape=(H, W))
                d = np.ctypeslib.as_array(C.cast(dst, C.POINTER(C.c_uint8)), shape=(H, W))
                e = cv2.createCLAHE().apply(a)
                d[:] = cv2.normalize(e, None, 0, 255, cv2.NORM_MINMAX)

            self._hooks = (LK_FN(lk), FM_FN(fm), EQ_FN(eq))
            lib().ora_tracker_set_hooks(self._h, *self._hooks)
        M = max(self.M, 1)
        self._bufs = {k: np.zeros(M, np.int32) for k in ("id", "track_cnt", "id_right")}
        self._bufs.update({k: np.zeros(M, np.float32) for k in
                           ("u", "v", "un_x", "un_y", "vx", "vy", "ru", "rv", "run_x", "run_y",
                            "rvx", "rvy")})
        self._t = _Tracks()
        for k, a in self._bufs.items():
            setattr(self._t, k, a.ctypes.data_as(_pi if a.dtype == np.int32 else _pf))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_tracker_destroy(self._h)
            self._h = None

    def track(self, cur_time, left, right, pub_this_frame=True, motion: dict | None = None):
        lx, ly, lt, lp = _ev(*left)
        rx, ry, rt, rp = _ev(*right)
        m = make_motion(motion) if motion is not None else None
        lib().ora_tracker_track_mc(self._h, float(cur_time), _p(lx), _p(ly), _p(lt), _p(lp), len(lx),
                                   _p(rx), _p(ry), _p(rt), _p(rp), len(rx), int(pub_this_frame),
                                   C.byref(m) if m is not None else None, C.byref(self._t))
        t = self._t
        nl, nr = t.n_left, t.n_right
        b = self._bufs
        out = {k: b[k][:nl].copy() for k in ("id", "track_cnt", "u", "v", "un_x", "un_y", "vx", "vy")}
        out.update({k: b[k][:nr].copy() for k in ("id_right", "ru", "rv", "run_x", "run_y", "rvx", "rvy")})
        out["stats"] = dict(n_prev=t.n_prev, n_after_temporal=t.n_after_temporal,
                            n_after_ransac=t.n_after_ransac, n_after_mask=t.n_after_mask,
                            n_new=t.n_new)
        return out

    def _unpack(self):
        t = self._t
        nl, nr = t.n_left, t.n_right
        b = self._bufs
        out = {k: b[k][:nl].copy() for k in ("id", "track_cnt", "u", "v", "un_x", "un_y", "vx", "vy")}
        out.update({k: b[k][:nr].copy() for k in ("id_right", "ru", "rv", "run_x", "run_y", "rvx", "rvy")})
        out["stats"] = dict(n_prev=t.n_prev, n_after_temporal=t.n_after_temporal,
                            n_after_ransac=t.n_after_ransac, n_after_mask=t.n_after_mask,
                            n_new=t.n_new)
        return out

    def track_image(self, cur_time, img_left, img_right=None, pub_this_frame=True):
        """FeatureTracker::trackImage (feature_tracker.cpp:164-338); cfg max_cnt / min_dist are
        MAX_CNT_IMG / MIN_DIST_IMG."""
        a = np.ascontiguousarray(img_left, np.uint8)
        assert a.shape == (self.H, self.W)
        b = None if img_right is None else np.ascontiguousarray(img_right, np.uint8)
        lib().ora_tracker_track_image(self._h, float(cur_time), _p(a),
                                      _p(b) if b is not None else None, int(pub_this_frame),
                                      C.byref(self._t))
        return self._unpack()

    def sae(self, cam):
        return Sae(self.W, self.H, _borrow=lib().ora_tracker_sae(self._h, cam))

    def time_surface(self, cam):
        p = lib().ora_tracker_time_surface(self._h, cam)
        return np.ctypeslib.as_array(p, shape=(self.H, self.W)).copy()

    def lk_image(self, cam):
        p = lib().ora_tracker_lk_image(self._h, cam)
        return np.ctypeslib.as_array(p, shape=(self.H, self.W)).copy()

    def next_id(self):
        """FeatureTracker::n_id (feature_tracker.cpp:9)."""
        return int(lib().ora_tracker_next_id(self._h))

    def timers(self):
        out = np.zeros(6, np.float64)
        lib().ora_tracker_timers(self._h, _p(out))
        return dict(zip(("sae", "ts", "lk_temporal", "select", "lk_stereo", "other"), out.tolist()))
