"""CPU suite: the node mirror pinned on the REFERENCE'S OWN event node.

oracle/_ref/libesvio_ref_node.so is feature_tracker/src/stereo_event_tracker_node.cpp compiled
UNMODIFIED (its main() renamed on the command line) on top of the unmodified feature_tracker.cpp
+ event_detector.cc (recipe: oracle/Makefile); ROS is a set of stand-in headers whose
Publisher::publish hands the message to the harness (oracle/ref_shim/ref_node_api.cc).

Each test plays one scripted stream through the reference's functions -- handle_stereo_event
(node.cpp:145-344), event_callback_left/right (:128-142), sync_process on its own thread
(:372-419), imu_callback / state_callback (:104-126) -- and through esvio_b200/node.py
(StereoEventNode, EventPairer, MotionAssembler) around the oracle tracker, and demands the same
decisions (publish gate, restarts, node state after every window, what the queues hold) and the
same published PointClouds bit for bit (header stamp, points, id*2+cam, u, v, vx, vy).  The C++
twin include/esvio_fe_node.hpp is held to node.py by the decision-trace test
(tests/test_node_logic.py), and the tracker inside to the reference's FeatureTracker by
tests/test_oracle_ref_tracker.py.
"""
import ctypes as C
import os
import time

import numpy as np
import pytest

from esvio_b200 import node, synth
from oracle import oracle as ora
from oracle import ref_tracker

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NODE_SO = os.path.join(ROOT, "oracle", "_ref", "libesvio_ref_node.so")
_p = ref_tracker._p
W, H = 346, 260


@pytest.fixture(scope="module")
def ref():
    if ref_tracker.load() is None or not os.path.exists(NODE_SO):   # load() runs `make ref` where it can
        pytest.skip("oracle/_ref/libesvio_ref_node.so not built and /root/reference absent")
    L = C.CDLL(NODE_SO)
    ev = [C.c_void_p] * 5 + [C.c_size_t]
    L.ref_node_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.ref_node_handle.argtypes = ev + ev + [C.c_uint32, C.c_uint32, C.c_double]
    L.ref_node_push_events.argtypes = [C.c_int] + ev + [C.c_uint32, C.c_uint32]
    L.ref_node_push_imu.argtypes = [C.c_double, C.c_void_p, C.c_void_p]
    L.ref_node_push_odometry.argtypes = [C.c_double, C.c_void_p]
    L.ref_node_queue_sizes.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_node_cloud.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_node_state.argtypes = [C.c_void_p] * 5
    L.ref_node_tracker_time.restype = C.c_double
    L.ref_node_tracker_prev_time.restype = C.c_double
    yield L
    L.ref_node_park_sync_thread()     # the detached sync_process thread must not run into interpreter shutdown


def _cfg_arrays(cfg):
    icfg = np.array([cfg["width"], cfg["height"], cfg["max_cnt"], cfg["min_dist"], cfg["flow_back"],
                     cfg["equalize"], cfg["ignore_polarity"], cfg["median_blur_kernel_size"],
                     int(cfg["focal_length"])], np.int32)
    d = [cfg["f_threshold"], cfg["ts_lk_threshold"], cfg["decay_ms"], cfg["feature_filter_threshold"]]
    for cam in cfg["cam"]:
        d += [cam[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")]
    return icfg, np.array(d, np.float64)


class _OracleFeatureTracker:
    """The oracle behind the reference's member names (what node.py drives)."""

    def __init__(self, cfg):
        self.t = ora.OracleTracker(cfg, use_cv2=False)
        self.PUB_THIS_FRAME = True
        cam = cfg["cam"][0]
        self.K = (cam["fx"], cam["fy"], cam["cx"], cam["cy"])
        self.cur_time = 0.0

    def trackEvent(self, cur_time, left, right, measurements=None):
        m = dict(measurements, K=self.K) if measurements is not None else None
        r = self.t.track(cur_time, left, right, self.PUB_THIS_FRAME, motion=m)
        self.cur_time = cur_time
        self.ids, self.track_cnt = r["id"], r["track_cnt"]
        self.cur_pts = np.stack([r["u"], r["v"]], 1)
        self.cur_un_pts = np.stack([r["un_x"], r["un_y"]], 1)
        self.pts_velocity = np.stack([r["vx"], r["vy"]], 1)
        self.ids_right = r["id_right"]
        self.cur_right_pts = np.stack([r["ru"], r["rv"]], 1)
        self.cur_un_right_pts = np.stack([r["run_x"], r["run_y"]], 1)
        self.right_pts_velocity = np.stack([r["rvx"], r["rvy"]], 1)


def _window(s, k, stamp_us, n=None):
    """(events of window k for both cameras incl. sec / nsec, header stamp (sec, nsec))."""
    L6, R6 = s.window(k, 0, n), s.window(k, 1, n)
    stamp = (synth.T0_SEC + stamp_us // 1_000_000, (stamp_us % 1_000_000) * 1000)
    return L6, R6, stamp


def _ev_args(e6):
    x, y, _, p, sec, nsec = (np.ascontiguousarray(a) for a in e6)
    return [_p(x), _p(y), _p(sec), _p(nsec), _p(p), len(x)], (x, y, p, sec, nsec)


def _msg(e6, stamp):
    return node.EventArray(node.to_sec(*stamp), e6[0], e6[1], e6[2], e6[3])


def _ref_state(L):
    ft, lt = C.c_double(), C.c_double()
    cnt, ff, ip = C.c_int(), C.c_int(), C.c_int()
    L.ref_node_state(C.byref(ft), C.byref(lt), C.byref(cnt), C.byref(ff), C.byref(ip))
    return ft.value, lt.value, cnt.value, bool(ff.value), bool(ip.value)


def _mirror_state(n):
    return n.first_image_time, n.last_image_time, n.pub_count, n.first_image_flag, n.init_pub


def _ref_clouds(L):
    out = []
    for i in range(L.ref_node_n_clouds()):
        n = L.ref_node_cloud_points(i)
        rows = np.zeros((n, 8), np.float32)
        sec, nsec = C.c_uint32(), C.c_uint32()
        L.ref_node_cloud(i, C.byref(sec), C.byref(nsec), _p(rows))
        out.append(((sec.value, nsec.value), rows))
    return out


def _same_clouds(ref_clouds, mirror_clouds):
    assert len(ref_clouds) == len(mirror_clouds)
    for (stamp, rows), c in zip(ref_clouds, mirror_clouds):
        assert stamp == node.ros_time(c.stamp)            # feature_points->header.stamp = ros::Time(msg_timestamp)
        assert rows.shape == c.rows.shape
        assert np.array_equal(rows.view(np.int32), np.ascontiguousarray(c.rows).view(np.int32))


@pytest.mark.parametrize("freq", [10, 15, 30])
def test_handle_stereo_event_equals_reference(ref, freq):
    """24 windows at 30 Hz with an empty left window, a forward jump of more than a second, a step
    back in time, and the publish-rate gate at FREQ 10 / 15 / 30 (config/*/es*io.yaml): same
    PUB_THIS_FRAME, same node state after every call, same restarts, same clouds."""
    cfg = synth.default_config(W, H)
    icfg, dcfg = _cfg_arrays(cfg)
    ref.ref_node_reset(_p(icfg), _p(dcfg), freq, 0)
    mt = _OracleFeatureTracker(cfg)
    mn = node.StereoEventNode(mt, freq)
    s = synth.StereoEventStream(W, H, 0.45e6)
    win_us = 1_000_000 // synth.WINDOWS_PER_SEC
    mirror_clouds, pubs = [], 0
    # script: (stream window, header stamp in us since T0, events per camera or None = all)
    script = [(k, (k + 1) * win_us, None) for k in range(10)]
    script.insert(4, (4, 5 * win_us, 0))                                 # an EventArray without events
    script += [(10, 11 * win_us + 1_500_000, None)]                      # > 1 s later: restart
    script += [(11 + k, (12 + k) * win_us + 1_500_000, None) for k in range(7)]
    script += [(18, 15 * win_us + 1_500_000, None)]                      # a step back in time: restart
    script += [(19 + k, (20 + k) * win_us + 1_500_000, None) for k in range(5)]
    for k, stamp_us, n in script:
        L6, R6, stamp = _window(s, k, stamp_us, n)
        la, keep_l = _ev_args(L6)
        ra, keep_r = _ev_args(R6)
        msg_t = node.to_sec(*stamp)                                       # sync_process: header stamp (:386,399)
        pub_ref = ref.ref_node_handle(*la, *ra, stamp[0], stamp[1], msg_t)
        tracked_before = mn.windows_tracked
        c = mn.handle_stereo_event(_msg(L6, stamp), _msg(R6, stamp), msg_t)
        if c is not None:
            mirror_clouds.append(c)
        assert _ref_state(ref) == _mirror_state(mn), (k, _ref_state(ref), _mirror_state(mn))
        if mn.windows_tracked > tracked_before:          # the call got as far as trackEvent
            assert bool(pub_ref) == bool(mt.PUB_THIS_FRAME), k
            assert ref.ref_node_tracker_time() == mt.cur_time == float(L6[2][-1])   # :190,193
        pubs += int(bool(pub_ref))
    assert ref.ref_node_n_restarts() == mn.restarts == 2
    _same_clouds(_ref_clouds(ref), mirror_clouds)
    assert len(mirror_clouds) >= 3 and sum(len(c.rows) for c in mirror_clouds) > 50
    assert 0 < pubs


def test_motion_compensation_assembly_equals_reference(ref):
    """Do_motion_correction = 1 (node.cpp:195-254): the Motion_correction_value assembled from the
    IMU / odometry queues -- one odometry message consumed per window, velocity-differenced
    acceleration, IMU messages before the first left event dropped, disordered IMU stamps refused
    -- and the clock / header stamp handed to trackEvent."""
    cfg = synth.default_config(W, H)
    icfg, dcfg = _cfg_arrays(cfg)
    ref.ref_node_reset(_p(icfg), _p(dcfg), 15, 1)
    ma = node.MotionAssembler()
    mt = _OracleFeatureTracker(cfg)
    mn = node.StereoEventNode(mt, 15, do_motion_correction=True, motion=ma)
    s = synth.StereoEventStream(W, H, 0.45e6, stream=3)
    win_us = 1_000_000 // synth.WINDOWS_PER_SEC
    mirror_clouds = []
    rng = np.random.default_rng(4)
    for k in range(12):
        L6, R6, stamp = _window(s, k, (k + 1) * win_us)
        t0 = float(L6[2][0])
        # IMU at 200 Hz-ish around the window, one of them out of order; odometry on most windows
        for j in range(5):
            t_imu = t0 - 0.004 + 0.006 * j if j != 3 else t0 - 0.010
            om = rng.normal(0, 1.2, 3)
            ac = rng.normal(0, 2.0, 3) + (0, 0, 9.805)
            ref.ref_node_push_imu(t_imu, _p(om), _p(ac))
            ma.push_imu(node.Imu(t_imu, tuple(om), tuple(ac)))
        if k % 4 != 1:
            v = np.array([0.3 * k, -0.2 * k * (k % 3), 0.05 * k * k])      # accelerations above and below 5 m/s^2
            ref.ref_node_push_odometry(t0 - 0.002, _p(v))
            ma.push_odometry(node.Odometry(t0 - 0.002, tuple(v)))
        la, _kl = _ev_args(L6)
        ra, _kr = _ev_args(R6)
        msg_t = node.to_sec(*stamp)
        ref.ref_node_handle(*la, *ra, stamp[0], stamp[1], msg_t)
        tracked_before = mn.windows_tracked
        c = mn.handle_stereo_event(_msg(L6, stamp), _msg(R6, stamp), msg_t)
        if c is not None:
            mirror_clouds.append(c)
        assert _ref_state(ref) == _mirror_state(mn), k
        if mn.windows_tracked > tracked_before:
            assert ref.ref_node_tracker_time() == mt.cur_time == float(L6[2][-1])   # :190,254
    _same_clouds(_ref_clouds(ref), mirror_clouds)
    assert len(mirror_clouds) >= 3


def _settle(L, timeout=20.0):
    """Wait until sync_process (2 ms poll, node.cpp:415) has taken what it can take: the queues
    stop changing and no trackEvent is running (prev_time == cur_time at its end, :587)."""
    deadline = time.time() + timeout
    stable, last = 0, None
    while time.time() < deadline:
        a, b = C.c_int(), C.c_int()
        L.ref_node_queue_sizes(C.byref(a), C.byref(b))
        cur = (a.value, b.value, L.ref_node_tracker_time(), L.ref_node_n_clouds(), _ref_state(L))
        busy = L.ref_node_tracker_prev_time() != L.ref_node_tracker_time()    # inside trackEvent
        stable = stable + 1 if (cur == last and not busy) else 0
        last = cur
        if stable >= 6:
            return a.value, b.value
        time.sleep(0.01)
    raise AssertionError("sync_process did not settle")


def test_pairing_through_the_callbacks_and_sync_process(ref):
    """Messages arrive through event_callback_left/right (depth-1 queues: a new message REPLACES
    the waiting one) while sync_process runs on its own thread, as in the node: pairs within the
    0.2 s tolerance are handled, an older side is thrown away, an overwritten message is never
    seen.  The mirror's EventPairer + StereoEventNode must end every step with the same queues,
    node state and clouds."""
    cfg = synth.default_config(W, H)
    icfg, dcfg = _cfg_arrays(cfg)
    ref.ref_node_reset(_p(icfg), _p(dcfg), 30, 0)
    ref.ref_node_start_sync_thread()
    mt = _OracleFeatureTracker(cfg)
    mn = node.StereoEventNode(mt, 30)
    pairer = node.EventPairer()
    s = synth.StereoEventStream(W, H, 0.3e6, stream=5)
    win_us = 1_000_000 // synth.WINDOWS_PER_SEC
    mirror_clouds = []

    def push(side, k, stamp_us):
        e6 = s.window(k, side)
        stamp = (synth.T0_SEC + stamp_us // 1_000_000, (stamp_us % 1_000_000) * 1000)
        args, _keep = _ev_args(e6)
        ref.ref_node_push_events(side, *args, stamp[0], stamp[1])
        (pairer.push_left if side == 0 else pairer.push_right)(_msg(e6, stamp))

    def step(pushes):
        for side, k, stamp_us in pushes:
            push(side, k, stamp_us)
        ql, qr = _settle(ref)
        while pairer.left and pairer.right:          # the consumer keeps up: poll until nothing more can happen
            pair = pairer.poll()
            if pair is not None:
                c = mn.handle_stereo_event(*pair)
                if c is not None:
                    mirror_clouds.append(c)
        assert (ql, qr) == (len(pairer.left), len(pairer.right))
        assert _ref_state(ref) == _mirror_state(mn)

    T = win_us
    step([(0, 0, 1 * T), (1, 0, 1 * T)])                          # a pair (first window: skipped by the node)
    step([(0, 1, 2 * T)])                                         # left alone: waits
    step([(1, 1, 2 * T)])                                         # its partner arrives
    step([(0, 2, 3 * T), (0, 3, 4 * T)])                          # the second left message replaces the first ...
    step([(1, 3, 4 * T)])                                         # ... and pairs with this one
    step([(1, 4, 5 * T)])                                         # right alone
    step([(0, 4, 5 * T + 150_000)])                               # 0.15 s apart: inside the tolerance
    step([(0, 5, 6 * T), (1, 5, 6 * T + 400_000)])                # left older than right - 0.2 s: left thrown away
    step([(0, 6, 6 * T + 450_000)])                               # pairs with the waiting right message
    step([(1, 7, 8 * T), (0, 7, 8 * T + 900_000)])                # right older than left - 0.2 s: right thrown away
    step([(1, 8, 8 * T + 900_000)])
    for k in range(9, 14):                                        # a run of ordinary pairs: clouds get published
        step([(0, k, (k + 20) * T), (1, k, (k + 20) * T)])
    assert ref.ref_node_n_restarts() == mn.restarts
    _same_clouds(_ref_clouds(ref), mirror_clouds)
    assert mn.windows_tracked >= 8 and len(mirror_clouds) >= 3


@pytest.mark.parametrize("frequency", [30.0, 1000.0])
def test_event_windower_equals_event_message_editor(ref, frequency):
    """The reference's re-packing tool (dependences/events_repacking_helper/src/
    EventMessageEditor.cpp:8-57, included unmodified) fed event by event, against
    node.EventWindower fed in chunks: same number of written messages, same header stamps (to the
    nanosecond, ros::Time rounding included), same events in each -- over a dense stretch, holes
    longer than a window (one short message per event until the end time has caught up), events
    exactly on a window's end, and repeated timestamps."""
    ref.ref_eme_create.restype = C.c_void_p
    ref.ref_eme_create.argtypes = [C.c_double]
    ref.ref_eme_insert.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_size_t]
    ref.ref_eme_count.argtypes = [C.c_void_p]
    ref.ref_eme_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    ref.ref_eme_destroy.argtypes = [C.c_void_p]
    rng = np.random.default_rng(7)
    period_us = int(round(1e6 / frequency))
    # event times in us since T0: dense noise, a hole of 3.7 windows, a burst of equal stamps, events
    # placed exactly on multiples of the window length after the first event, a long hole, a tail
    t0 = 123_456
    parts = [t0 + np.sort(rng.integers(0, 12 * period_us, 4000))]
    parts.append(parts[-1][-1] + int(3.7 * period_us) + np.sort(rng.integers(0, 5 * period_us, 1500)))
    parts.append(np.full(40, parts[-1][-1] + 17))
    parts.append(t0 + period_us * np.arange(30, 40))
    parts.append(parts[-1][-1] + 25 * period_us + np.sort(rng.integers(0, 6 * period_us, 2000)))
    us = np.sort(np.concatenate(parts)).astype(np.int64)
    sec = (synth.T0_SEC + us // 1_000_000).astype(np.uint32)
    nsec = ((us % 1_000_000) * 1000).astype(np.uint32)
    t = sec.astype(np.float64) + 1e-9 * nsec.astype(np.float64)
    n = len(us)
    x = rng.integers(0, W, n).astype(np.uint16)
    y = rng.integers(0, H, n).astype(np.uint16)
    p = rng.integers(0, 2, n).astype(np.uint8)
    h = ref.ref_eme_create(frequency)
    wd = node.EventWindower(frequency)
    msgs = []
    cuts = [0, 1, 2, 700, 701, 4000, 5533, n]             # chunk boundaries: arbitrary, incl. one-event chunks
    for a, b in zip(cuts, cuts[1:]):
        ref.ref_eme_insert(h, _p(x[a:b].copy()), _p(y[a:b].copy()), _p(sec[a:b].copy()), _p(nsec[a:b].copy()),
                           _p(p[a:b].copy()), b - a)
        msgs += wd.insert(x[a:b], y[a:b], t[a:b], p[a:b])
    assert ref.ref_eme_count(h) == len(msgs) > 40
    short = 0
    for i, m in enumerate(msgs):
        r = np.zeros(9, np.uint32)
        ref.ref_eme_get(h, i, _p(r))
        assert (int(r[0]), int(r[1])) == (int(r[2]), int(r[3])) == node.ros_time(m.stamp), i   # write time = header stamp
        assert int(r[4]) == len(m), i
        if len(m):
            assert node.to_sec(int(r[5]), int(r[6])) == m.t[0] and node.to_sec(int(r[7]), int(r[8])) == m.t[-1], i
        short += len(m) == 1
    assert short >= 3      # the holes produced their one-event messages
    ref.ref_eme_destroy(h)


# ---- the image node (stereo_image_tracker_node.cpp) -----------------------------------------------
IMG_SO = os.path.join(ROOT, "oracle", "_ref", "libesvio_ref_imgnode.so")
IW, IH = 240, 180


@pytest.fixture(scope="module")
def ref_img():
    if ref_tracker.load() is None or not os.path.exists(IMG_SO):
        pytest.skip("oracle/_ref/libesvio_ref_imgnode.so not built and /root/reference absent")
    L = C.CDLL(IMG_SO)
    L.ref_imgnode_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.ref_imgnode_handle.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
    L.ref_imgnode_push_image.argtypes = [C.c_int, C.c_void_p, C.c_double]
    L.ref_imgnode_queue_sizes.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_imgnode_cloud.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_imgnode_state.argtypes = [C.c_void_p] * 5
    L.ref_imgnode_tracker_time.restype = C.c_double
    L.ref_imgnode_tracker_prev_time.restype = C.c_double
    yield L
    L.ref_imgnode_park_sync_thread()


class _OracleImageTracker:
    """The oracle's trackImage behind the reference's member names; `clahe`: the image node's own
    CLAHE on both frames before the call (stereo_image_tracker_node.cpp:93-97)."""

    def __init__(self, cfg, clahe=False):
        self.t = ora.OracleTracker(cfg, use_cv2=False)
        self.PUB_THIS_FRAME = True
        self.clahe = clahe

    def trackImage(self, cur_time, left, right):
        if self.clahe:
            left, right = ora.clahe(left), ora.clahe(right)
        r = self.t.track_image(cur_time, left, right, self.PUB_THIS_FRAME)
        self.ids, self.track_cnt = r["id"], r["track_cnt"]
        self.cur_pts = np.stack([r["u"], r["v"]], 1)
        self.cur_un_pts = np.stack([r["un_x"], r["un_y"]], 1)
        self.pts_velocity = np.stack([r["vx"], r["vy"]], 1)
        self.ids_right = r["id_right"]
        self.cur_right_pts = np.stack([r["ru"], r["rv"]], 1)
        self.cur_un_right_pts = np.stack([r["run_x"], r["run_y"]], 1)
        self.right_pts_velocity = np.stack([r["rvx"], r["rvy"]], 1)


def _img_state(L):
    ft, lt = C.c_double(), C.c_double()
    cnt, ff, ip = C.c_int(), C.c_int(), C.c_int()
    L.ref_imgnode_state(C.byref(ft), C.byref(lt), C.byref(cnt), C.byref(ff), C.byref(ip))
    return ft.value, lt.value, cnt.value, bool(ff.value), bool(ip.value)


def _img_clouds(L):
    out = []
    for i in range(L.ref_imgnode_n_clouds()):
        rows = np.zeros((L.ref_imgnode_cloud_points(i), 8), np.float32)
        sec, nsec = C.c_uint32(), C.c_uint32()
        L.ref_imgnode_cloud(i, C.byref(sec), C.byref(nsec), _p(rows))
        out.append(((sec.value, nsec.value), rows))
    return out


@pytest.mark.parametrize("freq,equalize", [(10, 0), (20, 0), (10, 1)])
def test_handle_stereo_image_equals_reference(ref_img, freq, equalize):
    """handle_stereo_image (stereo_image_tracker_node.cpp:55-183) on a 14-frame stereo sequence at
    20 Hz with a jump of more than a second: the publish-rate gate, restart, first-publish
    suppression, the node's own CLAHE when EQUALIZE is set, and the published clouds."""
    cfg = synth.default_config(IW, IH, min_dist=14, max_cnt=60, equalize=equalize)
    icfg, dcfg = _cfg_arrays(cfg)
    ref_img.ref_imgnode_reset(_p(icfg), _p(dcfg), freq)
    mt = _OracleImageTracker(dict(cfg, equalize=0), clahe=bool(equalize))
    mn = node.StereoImageNode(mt, freq)
    frames = synth.stereo_frame_sequence(IW, IH, 14)
    mirror_clouds = []
    for k, (fl, fr_) in enumerate(frames):
        t = 1_700_000_000.0 + k / 20.0 + (1.6 if k >= 8 else 0.0)
        fl, fr_ = np.ascontiguousarray(fl), np.ascontiguousarray(fr_)
        pub_ref = ref_img.ref_imgnode_handle(_p(fl), _p(fr_), t)
        before = mn.windows_tracked
        c = mn.handle_stereo_image(fl, fr_, t)
        if c is not None:
            mirror_clouds.append(c)
        assert _img_state(ref_img) == _mirror_state(mn), k
        if mn.windows_tracked > before:
            assert bool(pub_ref) == bool(mt.PUB_THIS_FRAME), k
            assert ref_img.ref_imgnode_tracker_time() == t                     # :99 hands msg_timestamp
    assert ref_img.ref_imgnode_n_restarts() == mn.restarts == 1
    _same_clouds(_img_clouds(ref_img), mirror_clouds)
    assert len(mirror_clouds) >= 2 and sum(len(c.rows) for c in mirror_clouds) > 60


def test_image_pairing_through_the_callbacks_and_sync_process(ref_img):
    """mono8 frames through img_callback_left/right while the image node's sync_process runs on its
    own thread: depth-1 queues, the 1 s tolerance with its asymmetric comparisons (a left frame
    EXACTLY one second older than the right one is thrown, one exactly a second newer is not),
    getImageFromMsg.  ImagePairer + StereoImageNode must end every step with the same queues,
    node state and clouds."""
    cfg = synth.default_config(IW, IH, min_dist=14, max_cnt=60)
    icfg, dcfg = _cfg_arrays(cfg)
    ref_img.ref_imgnode_reset(_p(icfg), _p(dcfg), 20)
    ref_img.ref_imgnode_start_sync_thread()
    mt = _OracleImageTracker(cfg)
    mn = node.StereoImageNode(mt, 20)
    pairer = node.ImagePairer()
    frames = synth.stereo_frame_sequence(IW, IH, 14)
    mirror_clouds = []
    T0 = 1_700_000_000.0

    def settle():
        deadline, stable, last = time.time() + 20.0, 0, None
        while time.time() < deadline:
            a, b = C.c_int(), C.c_int()
            ref_img.ref_imgnode_queue_sizes(C.byref(a), C.byref(b))
            cur = (a.value, b.value, ref_img.ref_imgnode_tracker_time(), ref_img.ref_imgnode_n_clouds(),
                   _img_state(ref_img))
            busy = ref_img.ref_imgnode_tracker_prev_time() != ref_img.ref_imgnode_tracker_time()
            stable = stable + 1 if (cur == last and not busy) else 0
            last = cur
            if stable >= 6:
                return a.value, b.value
            time.sleep(0.01)
        raise AssertionError("sync_process did not settle")

    def step(pushes):
        for side, k, t in pushes:
            img = np.ascontiguousarray(frames[k][side])
            ref_img.ref_imgnode_push_image(side, _p(img), T0 + t)
            (pairer.push_left if side == 0 else pairer.push_right)(node.ImageMsg(T0 + t, img))
        ql, qr = settle()
        while pairer.left and pairer.right:
            pair = pairer.poll()
            if pair is not None:
                c = mn.handle_stereo_image(pair[0].image, pair[1].image, pair[2])
                if c is not None:
                    mirror_clouds.append(c)
        assert (ql, qr) == (len(pairer.left), len(pairer.right))
        assert _img_state(ref_img) == _mirror_state(mn)

    step([(0, 0, 0.00), (1, 0, 0.00)])                 # first pair: skipped by the node
    step([(0, 1, 0.05)])
    step([(1, 1, 0.05)])
    step([(0, 2, 0.10), (0, 3, 0.15)])                 # the waiting left frame is replaced
    step([(1, 3, 0.15)])
    step([(0, 4, 0.20), (1, 4, 1.20)])                 # left exactly 1 s older: thrown (`<=`)
    step([(0, 5, 2.20)])                               # left exactly 1 s newer than the waiting right: a pair (`>`);
    #                                                    2 s after the last handled frame: restart
    step([(0, 6, 2.25), (1, 6, 2.25)])                 # first frame after the restart
    step([(1, 7, 2.30), (0, 7, 3.35)])                 # right more than 1 s older: thrown
    step([(1, 8, 3.35)])                               # more than 1 s after 2.25: another restart
    for k in range(9, 14):
        step([(0, k, 3.0 + k * 0.05), (1, k, 3.0 + k * 0.05)])
    assert ref_img.ref_imgnode_n_restarts() == mn.restarts >= 1
    _same_clouds(_img_clouds(ref_img), mirror_clouds)
    assert mn.windows_tracked >= 5 and len(mirror_clouds) >= 1
