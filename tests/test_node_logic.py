"""CPU suite: the ROS-free mirror of the stereo_event_tracker node.
The C++ header (include/esvio_fe_node.hpp, driven by tests/cpp/node_logic.cpp with a stand-in
tracker) and the Python twin (esvio_b200/node.py) replay the same scripted streams and must
print the same trace: window boundaries of the fixed-rate re-windowing, left/right pairing,
first-window skip, publish-rate gate, motion measurements, PointCloud rows, first-publish
suppression, discontinuity restart."""
import os
import subprocess

import numpy as np
import pytest

from esvio_b200 import node

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "node_logic")


def _cpp_trace(arg=None):
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "node_logic.cpp"), "-o", EXE])
    r = subprocess.run([EXE] + ([arg] if arg else []), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    return r.stdout.splitlines()


def _stream(sec0, us0, n, step_us, gap_at, gap_us):
    i = np.arange(n, dtype=np.int64)
    us = us0 + i * step_us + np.where(i >= gap_at, gap_us, 0)
    sec = (sec0 + us // 1000000).astype(np.uint32)
    nsec = ((us % 1000000) * 1000).astype(np.uint32)
    t = sec.astype(np.float64) + 1e-9 * nsec.astype(np.float64)
    return ((i * 7) % 346).astype(np.uint16), ((i * 13) % 260).astype(np.uint16), t, (i & 1).astype(np.uint8)


class StubTracker:
    def __init__(self, log):
        self.PUB_THIS_FRAME = False
        self.calls = 0
        self.log = log

    def trackImage(self, t, img_left, img_right):
        self.trackEvent(t, ([0] * img_left,), ([0] * img_right,))

    def trackEvent(self, t, left, right, measurements=None):
        if measurements is not None:
            m = measurements
            self.log.append("M a=%.6f,%.6f,%.6f w=%.6f,%.6f,%.6f v=%.6f vp=%.6f t1=%.9f" % (
                *m["accel"], *m["omega"], m["state_v"][0], m["v_pre"][0], m["t1"]))
        nl, nr = len(left[0]), len(right[0])
        self.calls += 1
        c = self.calls
        n = 3 + c % 4
        self.ids = np.array([c + i for i in range(n)])
        self.track_cnt = np.array([1 + (i + c) % 3 for i in range(n)])
        self.cur_pts = np.array([[i, nl % 100] for i in range(n)], np.float32)
        self.cur_un_pts = np.array([[np.float32(0.1) * i, np.float32(0.2) * i] for i in range(n)], np.float32)
        self.pts_velocity = np.array([[1, 2]] * n, np.float32)
        ev = [i for i in range(n) if i % 2 == 0]
        self.ids_right = np.array([c + i for i in ev])
        self.cur_right_pts = np.array([[i - 5.0, nr % 100] for i in ev], np.float32).reshape(-1, 2)
        self.cur_un_right_pts = np.zeros((len(ev), 2), np.float32)
        self.right_pts_velocity = np.array([[1.5, 2.5]] * len(ev), np.float32).reshape(-1, 2)
        self.log.append("T %.9f nl=%d nr=%d pub=%d" % (t, nl, nr, int(self.PUB_THIS_FRAME)))


def _py_trace(mc):
    log = []
    lm = node.window_stream(_stream(1700000000, 100, 60000, 25, 30000, 1400000), 30.0, chunk=7777)
    rm = node.window_stream(_stream(1700000000, 3100, 59970, 25, 30000, 1400000), 30.0)
    log += ["WL %.9f %d" % (m.stamp, len(m)) for m in lm]
    log += ["WR %.9f %d" % (m.stamp, len(m)) for m in rm]
    trk = StubTracker(log)
    nd = node.StereoEventNode(trk, 15, do_motion_correction=mc)
    if mc:
        for i in range(400):
            nd.motion.push_imu(node.Imu(1700000000.0 + 0.005 * i, (0.01 * i, -0.02 * i, 0.5), (0, 0, 0)))
        assert not nd.motion.push_imu(node.Imu(1700000000.0, (9, 9, 9), (0, 0, 0)))
        for i in range(40):
            nd.motion.push_odometry(node.Odometry(1700000000.0 + 0.05 * i, (0.1 * i * i, 0.2, -0.1 * i)))

    def handle(l, r, ts, with_restarts=False):
        c = nd.handle_stereo_event(l, r, ts)
        line = "H %.9f published=%d rows=%d" % (ts, int(c is not None), len(c.rows) if c is not None else 0)
        if with_restarts:
            line += " restarts=%d" % nd.restarts
        elif c is not None:
            line += "".join(" %g:%g" % (row[3], row[4]) for row in c.rows)
        log.append(line)

    pairer = node.EventPairer()
    il = ir = 0
    while il < len(lm) or ir < len(rm):
        left = ir >= len(rm) or (il < len(lm) and lm[il].stamp <= rm[ir].stamp)
        if left:
            pairer.push_left(lm[il]); il += 1
        else:
            pairer.push_right(rm[ir]); ir += 1
        while pairer.left and pairer.right:
            pair = pairer.poll()
            if pair is not None:
                handle(*pair)
    l, r = lm[-1], rm[-1]
    base = l.stamp
    for ts in (base + 2.0, base + 2.033, base + 2.066, base + 2.0):
        handle(l, r, ts, with_restarts=True)
    empty = node.EventArray(0.0, *(np.zeros(0, d) for d in (np.uint16, np.uint16, np.float64, np.uint8)))
    log.append("E %d" % int(nd.handle_stereo_event(empty, r, base + 3.0) is not None))
    log.append("END dropped=%d restarts=%d tracked=%d" % (pairer.dropped, nd.restarts, nd.windows_tracked))
    return log


@pytest.mark.parametrize("mc", [False, True])
def test_cpp_and_python_node_agree(mc):
    cpp = _cpp_trace("m" if mc else None)
    py = _py_trace(mc)
    assert len(cpp) == len(py), (len(cpp), len(py))
    for a, b in zip(cpp, py):
        assert a == b, (a, b)
    # the script exercises every branch
    assert any(l.startswith("H") and "published=1" in l for l in py)
    assert py[-1].startswith("END dropped=0 restarts=2")
    hs = [l for l in py if l.startswith("H ")]
    assert "published=0" in hs[0]                       # first pair only arms the node
    assert sum("published=1" in l for l in hs) >= 20    # ~15 Hz out of 30 Hz windows
    if mc:
        assert sum(l.startswith("M ") for l in py) == sum(l.startswith("T ") for l in py)


def test_windower_matches_event_message_editor_rules():
    """EventMessageEditor.cpp:34-50: first event opens the window; an event at/after the end
    flushes (stamp = end, ns-rounded) and reopens AT the end; after a hole, one message per
    event until the end time catches up."""
    x, y, t, p = _stream(1700000000, 0, 100, 1000, 50, 200000)   # 1 kHz, 0.2 s hole after 50
    msgs = node.window_stream((x, y, t, p), 30.0)
    assert len(msgs[0]) == 34 and abs(msgs[0].stamp - (t[0] + 1 / 30)) < 1e-6
    assert len(msgs[1]) == 16                       # events 34..49, flushed by event 50
    tail = [len(m) for m in msgs[2:8]]
    assert tail[:5] == [1] * 5 and tail[5] > 1      # hole: one-event messages until caught up
    assert all(abs((b.stamp - a.stamp) - 1 / 30) < 1e-6 for a, b in zip(msgs, msgs[1:]))
    total = sum(len(m) for m in msgs)
    assert total < 100                              # the open last window is never written


def test_pairer_tolerance_and_depth_one_queues():
    z = [np.zeros(1, d) for d in (np.uint16, np.uint16, np.float64, np.uint8)]
    mk = lambda s: node.EventArray(s, *z)
    p = node.EventPairer()
    p.push_left(mk(10.0)); p.push_left(mk(10.033))      # second replaces the first
    assert p.dropped == 1 and p.poll() is None
    p.push_right(mk(10.5))                              # left is older than right - 0.2: dropped
    assert p.poll() is None and not p.left and p.right
    p.push_left(mk(10.45))
    l, r, ts = p.poll()
    assert ts == 10.45 and l.stamp == 10.45 and r.stamp == 10.5
    p.push_left(mk(11.0)); p.push_right(mk(10.7))       # right too old: dropped
    assert p.poll() is None and p.left and not p.right


def test_cloud_rows_decode_like_the_estimator():
    """stereo_estimator_node.cpp:388-401 + feature_manager.cpp:331-340: per id the first row is
    camera 0 and at most one more row, camera 1."""
    log = []
    t = StubTracker(log)
    t.trackEvent(1.0, (np.zeros(5),), (np.zeros(5),))
    rows = node.pack_feature_cloud(t)
    fid, cam = node.decode_feature_cloud(rows)
    assert set(fid[cam == 1]) <= set(fid[cam == 0])
    assert (np.diff(np.flatnonzero(cam == 0)) == 1).all() and cam[0] == 0
    assert (np.asarray(t.track_cnt)[np.isin(t.ids, fid[cam == 0])] > 1).all()


def _py_image_trace():
    log = []
    trk = StubTracker(log)
    nd = node.StereoImageNode(trk, 10)
    pairer = node.ImagePairer()
    t0 = 1700000000.0

    def step(left, stamp, tag):
        m = node.ImageMsg(stamp, tag)           # the stand-in "image" is its tag
        (pairer.push_left if left else pairer.push_right)(m)
        while pairer.left and pairer.right:
            pair = pairer.poll()
            if pair is None:
                log.append("D")
                continue
            l, r, ts = pair
            c = nd.handle_stereo_image(l.image, r.image, ts)
            line = "H %.9f l=%d r=%d published=%d rows=%d restarts=%d" % (
                ts, l.image, r.image, int(c is not None), len(c.rows) if c is not None else 0, nd.restarts)
            if c is not None:
                line += "".join(" %g:%g" % (row[3], row[4]) for row in c.rows)
            log.append(line)

    for k in range(60):
        tl = t0 + 0.05 * k + (1.5 if k > 40 else 0.0)
        tr = tl + 0.002
        if k == 30:
            tr = tl + 1.0
        if k == 59:
            tl = tr = t0 + 0.05 * 50
        step(True, tl, 1000 + k)
        if k != 7:
            step(False, tr, 2000 + k)
    log.append("END dropped=%d restarts=%d tracked=%d" % (pairer.dropped, nd.restarts, nd.windows_tracked))
    return log


def test_cpp_and_python_image_node_agree():
    """handle_stereo_image + the image node's pairing step (stereo_image_tracker_node.cpp:54-183,
    217-241): the C++ header and the Python twin print the same decision trace on a scripted
    stream with a lost right frame, a left frame exactly 1 s older than the right one (thrown by
    the image node's `<=`), a 1.5 s hole (restart) and a frame that goes back in time."""
    cpp = [l for l in _cpp_trace("i") if not l.startswith("T ")]
    py = [l for l in _py_image_trace() if not l.startswith("T ")]
    assert len(cpp) == len(py), (len(cpp), len(py), cpp[:5], py[:5])
    for a, b in zip(cpp, py):
        assert a == b, (a, b)
    assert sum(l == "D" for l in py) >= 1                     # the `<=` throw happened
    assert py[-1].startswith("END") and "restarts=2" in py[-1]
    hs = [l for l in py if l.startswith("H ")]
    assert "published=0" in hs[0] and sum("published=1" in l for l in hs) >= 15


def test_replay_frames_host_logic_with_a_stub_tracker():
    """esvio_b200.replay --frames without a GPU: the message builder and replay_images around a
    stub tracker (every frame reaches trackImage once, in order; first pair only arms the node)."""
    import argparse
    from esvio_b200 import replay
    args = argparse.Namespace(workload="stereo_davis346_1mevs", npz=None, windows=12)
    W, H, freq, lm, rm = replay.frame_messages(args)
    assert (W, H, freq) == (346, 260, 15) and len(lm) == len(rm) == 12
    assert lm[0].image.shape == (260, 346) and lm[0].image.dtype == np.uint8

    class Stub(StubTracker):
        def trackImage(self, t, img_left, img_right):
            assert img_left.shape == (260, 346) and img_right.shape == (260, 346)
            self.trackEvent(t, ([0] * 7,), ([0] * 5,))

    log = []
    nd = node.StereoImageNode(Stub(log), freq)
    clouds, dropped = node.replay_images(nd, lm, rm)
    assert dropped == 0 and nd.windows_tracked == 11 and nd.restarts == 0
    stamps = [float(l.split()[1]) for l in log if l.startswith("T ")]
    assert stamps == sorted(stamps) and len(stamps) == 11
    assert 4 <= len(clouds) <= 10 and all(c.rows.shape[1] == 8 for c in clouds)


@pytest.mark.parametrize("mc", [False, True])
def test_track_event_clock_is_the_last_left_event(mc):
    """Written from the reference, not from the other twin: stereo_event_tracker_node.cpp:190
    `msg_timestamp_left = event_left.events.back().ts.toSec()` is what BOTH trackEvent calls get
    (:193 and :254); the header stamp only travels as t_left_1 inside the measurements
    (:197,202) and as the cloud's stamp (:280)."""
    log = []
    trk = StubTracker(log)
    nd = node.StereoEventNode(trk, 15, do_motion_correction=mc)
    if mc:
        nd.motion.push_imu(node.Imu(1700000000.0, (0.1, 0.2, 0.3), (0, 0, 0)))
    x, y, t, p = _stream(1700000000, 100, 4000, 25, 10 ** 9, 0)
    stamps, lasts = [], []
    for k in range(4):
        sl = slice(1000 * k, 1000 * (k + 1))
        stamp = float(t[sl][-1]) + 0.004            # header stamp = window end, after the last event
        msg = node.EventArray(stamp, x[sl], y[sl], t[sl], p[sl])
        nd.handle_stereo_event(msg, msg, stamp)
        if k:                                       # the first pair only arms the node
            stamps.append(stamp)
            lasts.append(float(t[sl][-1]))
    times = [float(l.split()[1]) for l in log if l.startswith("T ")]
    assert times == pytest.approx(lasts, abs=1e-9) and len(times) == 3
    assert all(abs(a - b) > 1e-3 for a, b in zip(times, stamps))
    if mc:
        t1 = [float(l.split("t1=")[1]) for l in log if l.startswith("M ")]
        assert t1 == pytest.approx(stamps, abs=1e-9)


def test_cpp_node_hands_the_same_clock_with_and_without_motion_compensation():
    """The C++ twin on the same script: the `T <time>` lines (the clock trackEvent got) must not
    depend on Do_motion_correction (node.cpp:190 feeds both :193 and :254)."""
    plain = [l.split()[1] for l in _cpp_trace(None) if l.startswith("T ")]
    mc = [l.split()[1] for l in _cpp_trace("m") if l.startswith("T ")]
    assert plain == mc and len(plain) > 20
