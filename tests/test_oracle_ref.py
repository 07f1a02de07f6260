"""CPU suite: the oracle pinned on the REFERENCE'S OWN CODE.

oracle/_ref/libesvio_ref.so is the reference's feature_tracker/src/event_detector/
event_detector.cc, compiled unmodified from /root/reference against the stand-in headers in
oracle/ref_shim/ (recipe: oracle/Makefile).  Every test drives the same seeded synthetic event
streams through that library and through the oracle's restatement and demands identical
results -- bit-exact SAE planes, identical CV_8U time surfaces, identical Arc* decisions -- for
createSAE_left/right (event_detector.cc:149-166,212-228), SAEtoTimeSurface_left/right
(:230-305), isCorner (:308-544) and the motion-compensated createSAE_* (:102-147,168-210 with
motioncorrection :547-591; there the shim's Matrix3f arithmetic is the oracle's statement of
Eigen's kernels, so only the reference's control flow is pinned -- see mini_eigen.h).

The GPU parity tests compare the CUDA path with the oracle bit for bit on the same kinds of
streams, which closes the chain  reference code == oracle == CUDA  for these stages.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from esvio_b200 import synth
from oracle import oracle as ora

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libesvio_ref.so")
REF_SRC = "/root/reference/feature_tracker/src/event_detector/event_detector.cc"

_p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
_EXP_HOOK = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.POINTER(C.c_float))


@pytest.fixture(scope="module")
def ref():
    if os.path.exists(REF_SRC):  # this container: (re)build from the reference where it lies
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref"])
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libesvio_ref.so not built and /root/reference absent")
    L = C.CDLL(REF_SO)
    L.ref_create.restype = C.c_void_p
    L.ref_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int,
                             C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
    L.ref_destroy.argtypes = [C.c_void_p]
    L.ref_update.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_size_t]
    L.ref_update_mc.argtypes = ([C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_size_t]
                                + [C.c_void_p] * 4 + [C.c_double, C.c_double])
    L.ref_time_surface.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
    L.ref_corner_flags.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_size_t, C.c_void_p]
    L.ref_get_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.ref_set_exp_hook.argtypes = [_EXP_HOOK]
    return L


class RefDetector:
    """esvio::EventDetector of the reference behind oracle/ref_shim/ref_api.cc."""

    def __init__(self, L, W, H, decay_ms=20.0, ignore_polarity=0, median_k=0, filter_threshold=0.01,
                 min_dist=10, K=None):
        self.L, self.W, self.H = L, W, H
        k = K or (0.0, 0.0, 0.0, 0.0)
        self.h = L.ref_create(W, H, decay_ms, ignore_polarity, median_k, filter_threshold, min_dist,
                              int(K is not None), *[float(v) for v in k])

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_destroy(self.h)
            self.h = None

    @staticmethod
    def _ev(x, y, t, p):
        return (np.ascontiguousarray(x, np.uint16), np.ascontiguousarray(y, np.uint16),
                np.ascontiguousarray(t, np.float64), np.ascontiguousarray(p, np.uint8))

    def update(self, cam, x, y, t, p):
        x, y, t, p = self._ev(x, y, t, p)
        self.L.ref_update(self.h, cam, _p(x), _p(y), _p(t), _p(p), len(x))

    def update_mc(self, cam, x, y, t, p, use_mc, m, t0, t1):
        x, y, t, p = self._ev(x, y, t, p)
        use = np.ascontiguousarray(use_mc, np.uint8)
        st = np.array(list(m["state_v"]) + [0.0], np.float64)
        vp, a, w = (np.array(m[k], np.float32) for k in ("v_pre", "accel", "omega"))
        self.L.ref_update_mc(self.h, cam, _p(x), _p(y), _p(t), _p(p), _p(use), len(x), _p(st), _p(vp),
                             _p(a), _p(w), float(t0), float(t1))

    def planes(self, cam):
        out = []
        for which in (0, 1):
            for pol in (0, 1):
                a = np.empty((self.H, self.W), np.float64)
                self.L.ref_get_plane(self.h, cam, which, pol, _p(a))
                out.append(a)
        return out   # sae[0], sae[1], latest[0], latest[1] -- the order of oracle.Sae.planes()

    def time_surface(self, cam, t_ref):
        out = np.empty((self.H, self.W), np.uint8)
        self.L.ref_time_surface(self.h, cam, t_ref, _p(out))
        return out

    def corner_flags(self, x, y, t, p):
        x, y, t, p = self._ev(x, y, t, p)
        out = np.zeros(len(x), np.uint8)
        self.L.ref_corner_flags(self.h, _p(x), _p(y), _p(t), _p(p), len(x), _p(out))
        return out


def _same_planes(a, b):
    for k, (u, v) in enumerate(zip(a, b)):
        assert np.array_equal(u, v), f"plane {k}: {np.count_nonzero(u != v)} px differ"


@pytest.mark.parametrize("W,H,rate,windows,min_dist", [
    (346, 260, 1.0e6, 4, 10),       # BASELINE configs[1]
    (640, 480, 5.0e6, 2, 10),       # configs[2]
    (640, 480, 5.0e6, 1, 30),       # config/esvio_DSEC (min_dist decides isCorner's border)
    (640, 480, 20.0e6, 1, 20),      # configs[3] burst rate, config/esvio_ecmd min_dist
])
def test_sae_time_surface_arcstar_equal_the_reference(ref, W, H, rate, windows, min_dist):
    s = synth.StereoEventStream(W, H, rate)
    det = RefDetector(ref, W, H, min_dist=min_dist)
    o = [ora.Sae(W, H), ora.Sae(W, H)]
    n_corner = 0
    for k in range(windows):
        L, R, t_ref = s.stereo_window(k)
        for cam, ev in enumerate((L, R)):
            det.update(cam, *ev)
            o[cam].update(*ev)
        for cam in range(2):
            _same_planes(det.planes(cam), o[cam].planes())
            assert np.array_equal(det.time_surface(cam, t_ref), o[cam].time_surface(t_ref))
        # the reference tests corners after the whole window is inserted (feature_tracker.cpp:356-362
        # precede :458), on the left camera's planes
        f_ref = det.corner_flags(*L)
        f_ora = o[0].corner_flags(*L, min_dist=min_dist)
        assert np.array_equal(f_ref, f_ora), np.count_nonzero(f_ref != f_ora)
        n_corner += int(f_ref.sum())
    assert n_corner > 50          # the comparison saw real corners, not all-zero flags


def test_edge_cases_equal_the_reference(ref):
    """Same-pixel / same-timestamp storm, polarity flips inside and outside the refractory
    window, empty input, and a reference time far in the future (everything decayed)."""
    W, H = 346, 260
    rng = np.random.default_rng(7)
    det, o = RefDetector(ref, W, H), ora.Sae(W, H)
    n = 50_000
    x = rng.integers(100, 104, n).astype(np.uint16)          # 16 pixels take all events
    y = rng.integers(50, 54, n).astype(np.uint16)
    t = 1.7e9 + np.sort(rng.integers(0, 30_000, n)) * 1e-6   # many equal timestamps
    p = rng.integers(0, 2, n).astype(np.uint8)
    for ev in ((x, y, t, p), tuple(a[:0] for a in (x, y, t, p))):
        det.update(0, *ev)
        o.update(*ev)
        _same_planes(det.planes(0), o.planes())
    for t_ref in (float(t[-1]), float(t[-1]) + 0.02, float(t[-1]) + 5.0):
        assert np.array_equal(det.time_surface(0, t_ref), o.time_surface(t_ref))
    untouched = det.time_surface(0, float(t[-1]))[0, 0]
    assert untouched == 128                                   # 127.5 -> 128, round half to even


@pytest.mark.parametrize("ignore_polarity,median_k", [(1, 0), (0, 1), (1, 2)])
def test_time_surface_options_equal_the_reference(ref, ignore_polarity, median_k):
    """ignore_polarity (event_detector.cc:246-259) and the median blur (:262-264)."""
    W, H = 346, 260
    s = synth.StereoEventStream(W, H, 1.0e6)
    det = RefDetector(ref, W, H, ignore_polarity=ignore_polarity, median_k=median_k)
    o = ora.Sae(W, H)
    for k in range(2):
        L, _, t_ref = s.stereo_window(k)
        det.update(0, *L)
        o.update(*L)
    ts = o.time_surface(t_ref, ignore_polarity=ignore_polarity)
    if median_k:
        ts = ora.median_blur(ts, 2 * median_k + 1)
    assert np.array_equal(det.time_surface(0, t_ref), ts)


@pytest.mark.parametrize("omega,accel", [
    ((0.4, -0.3, 0.8), (6.0, 1.0, -2.0)),       # Pade-3 branch of Matrix3f::exp()
    ((25.0, -14.0, 31.0), (0.0, 7.5, 0.0)),     # larger rotation: Pade-5
    ((0.4, -0.3, 0.8), (1.0, 1.0, 1.0)),        # |a| <= 5: the warp is gated off (:125)
])
def test_motion_compensated_update_equals_the_reference(ref, omega, accel):
    """createSAE_*(…, measurements) + motioncorrection as the reference wrote them; the per-event
    choice of overload is trackEvent's rule (feature_tracker.cpp:628-642): warp iff dt > 0 and
    (t - t0) / dt < 1 with t0 = first left event, t1 = left header stamp."""
    W, H = 346, 260
    K = (250.0, 249.0, 173.0, 130.0)
    hook = _EXP_HOOK(lambda a, o: ora.lib().ora_mat3_exp_f(a, o))
    ref.ref_set_exp_hook(hook)
    s = synth.StereoEventStream(W, H, 1.0e6)
    det = RefDetector(ref, W, H, K=K)
    o = [ora.Sae(W, H), ora.Sae(W, H)]
    for k in range(2):
        L, R, t_ref = s.stereo_window(k)
        t0 = float(L[2][0])
        t1 = t0 + 0.8 * (float(L[2][-1]) - t0)     # header stamp inside the window: both overloads run
        m = dict(state_v=(0.8, -0.4, 0.2), v_pre=(0.7, -0.5, 0.1), accel=accel, omega=omega, t1=t1, K=K)
        for cam, ev in enumerate((L, R)):
            dt = t1 - t0
            use = (dt > 0) & ((ev[2] - t0) / dt < 1)
            assert 0 < use.sum() < len(use)
            det.update_mc(cam, *ev, use, m, t0, t1)
            o[cam].update_mc(*ev, m, t0)
            _same_planes(det.planes(cam), o[cam].planes())
            assert np.array_equal(det.time_surface(cam, t_ref), o[cam].time_surface(t_ref))
    if ora.motion_active(m):      # the warp moved events: the planes differ from the unwarped run
        plain = ora.Sae(W, H)
        for k in range(2):
            plain.update(*s.stereo_window(k)[0])
        assert any(not np.array_equal(a, b) for a, b in zip(plain.planes(), o[0].planes()))
