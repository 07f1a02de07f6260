"""ROS-free mirror of the reference's stereo_event_tracker node around the tracker call
(SURVEY.md section 8f rank 1): fixed-rate windowing of a raw stream, left/right pairing,
`handle_stereo_event` (first-window skip, discontinuity restart, publish-rate gate, motion
measurements, PointCloud packing, first-publish suppression) and a replay loop.

  EventWindower     dependences/events_repacking_helper/src/EventMessageEditor.cpp:8-57
  EventPairer       feature_tracker/src/stereo_event_tracker_node.cpp:128-142, 372-419
  MotionAssembler   stereo_event_tracker_node.cpp:102-125, 195-252
  StereoEventNode   stereo_event_tracker_node.cpp:145-344

The node only talks to a tracker object with the reference's member names (`trackEvent`,
`PUB_THIS_FRAME`, `ids`, `track_cnt`, `cur_pts`, ...): `esvio_b200.frontend.FeatureTracker` on
the GPU, or any stand-in with the same names in the host-logic tests.  Nothing here computes
on events -- that is libesvio_fe.so's job.
"""
from __future__ import annotations

import math
from collections import deque
from dataclasses import dataclass, field

import numpy as np


def ros_time(t: float):
    """ros::Time(double).fromSec: (sec, nsec) with nsec rounded to the nearest integer."""
    sec = int(math.floor(t))
    nsec = int(round((t - sec) * 1e9))
    sec += nsec // 1000000000
    nsec %= 1000000000
    return sec, nsec


def to_sec(sec: int, nsec: int) -> float:
    """ros::Time::toSec()"""
    return float(sec) + 1e-9 * float(nsec)


@dataclass
class EventArray:
    """dvs_msgs/EventArray: header stamp + SoA events (x, y, t, p); t = e.ts.toSec()."""
    stamp: float
    x: np.ndarray
    y: np.ndarray
    t: np.ndarray
    p: np.ndarray

    def __len__(self):
        return len(self.x)

    @property
    def events(self):
        return (self.x, self.y, self.t, self.p)


class EventWindower:
    """EventMessageEditor (EventMessageEditor.cpp:8-57): re-packs a raw event stream into
    EventArray messages of 1/frequency seconds.  The first event opens a window at its own
    time; an event at or after the window's end flushes the window (header stamp = the end
    time, rounded to ns like ros::Time) and opens the next one AT THAT END TIME -- so after a
    gap every event flushes one (short) message until the end time has caught up, exactly as
    the reference tool does."""

    def __init__(self, frequency: float = 30.0):
        self.duration = 1.0 / frequency
        self.first = True
        self.start = self.end = 0.0
        self._buf = []

    def _reset(self, start: float):
        self.start = start
        self.end = to_sec(*ros_time(start + self.duration))
        self._buf = []

    def insert(self, x, y, t, p):
        """Feeds a time-ascending chunk; returns the EventArray messages it completed."""
        out = []
        x, y, t, p = (np.asarray(a) for a in (x, y, t, p))
        i, n = 0, len(t)
        if n and self.first:
            self._reset(float(t[0]))
            self.first = False
        while i < n:
            # events before the current end time go to the open window in one slice
            j = i + int(np.searchsorted(t[i:], self.end, side="left"))
            if j > i:
                self._buf.append((x[i:j], y[i:j], t[i:j], p[i:j]))
                i = j
            if i < n:  # t[i] >= end: flush, reopen at the end time, then take this event
                out.append(self._flush())
                self._buf.append((x[i:i + 1], y[i:i + 1], t[i:i + 1], p[i:i + 1]))
                i += 1
        return out

    def _flush(self):
        parts = self._buf
        if parts:
            arrs = [np.concatenate([q[k] for q in parts]) for k in range(4)]
        else:
            arrs = [np.zeros(0, np.uint16), np.zeros(0, np.uint16), np.zeros(0, np.float64),
                    np.zeros(0, np.uint8)]
        msg = EventArray(self.end, *arrs)
        self._reset(self.end)
        return msg


class EventPairer:
    """The depth-1 queues of event_callback_left/right (node.cpp:128-142: a new message
    REPLACES the waiting one) and the pairing step of sync_process (node.cpp:380-406)."""

    def __init__(self, tolerance: float = 0.2):
        self.tol = tolerance
        self.left = deque()
        self.right = deque()
        self.dropped = 0

    def push_left(self, msg: EventArray):
        if self.left:
            self.left.popleft()
            self.dropped += 1
        self.left.append(msg)

    def push_right(self, msg: EventArray):
        if self.right:
            self.right.popleft()
            self.dropped += 1
        self.right.append(msg)

    def poll(self):
        """One iteration of sync_process: (left, right, msg_timestamp) or None."""
        if not self.left or not self.right:
            return None
        tl, tr = self.left[0].stamp, self.right[0].stamp
        if tl < tr - self.tol:
            self.left.popleft()
            return None
        if tl > tr + self.tol:
            self.right.popleft()
            return None
        l, r = self.left.popleft(), self.right.popleft()
        return l, r, l.stamp


@dataclass
class Imu:
    stamp: float
    angular_velocity: tuple
    linear_acceleration: tuple


@dataclass
class Odometry:
    stamp: float
    linear_velocity: tuple


class MotionAssembler:
    """Builds the Motion_correction_value of one window from the IMU / odometry queues
    (node.cpp:195-252), including its quirks: odometry is consumed one message per window,
    the acceleration handed to the tracker is the velocity-differenced `temp_a` in float,
    IMU messages older than the first left event are discarded and the first remaining one
    supplies omega."""

    def __init__(self):
        self.imu = deque()
        self.odom = deque()
        self.last_imu_t = 0.0
        self.v_cur = np.zeros(3, np.float32)   # Eigen::Vector3f globals (node.cpp:53-54)
        self.v_pre = np.zeros(3, np.float32)
        self.t_cur = 0.0
        self.t_pre = 0.0
        self.is_nolinear = False

    def push_imu(self, m: Imu):
        if m.stamp <= self.last_imu_t:   # "imu message in disorder!" (node.cpp:111-115)
            return False
        self.last_imu_t = m.stamp
        self.imu.append(m)
        return True

    def push_odometry(self, m: Odometry):
        self.odom.append(m)
        self.is_nolinear = True

    def assemble(self, t_left_0: float, t_left_1: float) -> dict:
        state_v = np.zeros(3)
        temp_a = np.zeros(3, np.float32)
        omega = np.zeros(3, np.float32)
        if self.imu:
            if self.odom:
                o = self.odom.popleft()
                state_v = np.asarray(o.linear_velocity, np.float64)
                self.v_pre = self.v_cur.copy()
                self.v_cur = state_v.astype(np.float32)
                self.t_pre, self.t_cur = self.t_cur, o.stamp
                with np.errstate(divide="ignore", invalid="ignore"):
                    # float difference, divided in double, stored as float (node.cpp:230-232)
                    temp_a = ((self.v_cur - self.v_pre).astype(np.float64)
                              / (self.t_cur - self.t_pre)).astype(np.float32)
            while self.imu and self.imu[0].stamp < t_left_0:
                self.imu.popleft()
            if self.imu:
                omega = np.asarray(self.imu[0].angular_velocity, np.float32)
        return dict(state_v=tuple(state_v), v_pre=tuple(self.v_pre), accel=tuple(temp_a),
                    omega=tuple(omega), t1=t_left_1)


@dataclass
class FeatureCloud:
    """sensor_msgs/PointCloud as the node fills it (node.cpp:276-331): frame_id "world",
    points (x, y, 1), channels id*2+cam, u, v, vx, vy."""
    stamp: float
    rows: np.ndarray = field(default_factory=lambda: np.zeros((0, 8), np.float32))


class StereoEventNode:
    """handle_stereo_event (node.cpp:145-344) around any tracker with the reference's names."""

    NUM_OF_CAM_stereo = 2

    def __init__(self, tracker, freq: int, do_motion_correction: bool = False,
                 motion: MotionAssembler | None = None):
        self.t = tracker
        self.FREQ = freq
        self.do_mc = do_motion_correction
        self.motion = motion or MotionAssembler()
        self.first_image_flag = True
        self.first_image_time = 0.0
        self.last_image_time = 0.0
        self.pub_count = 1
        self.init_pub = False
        self.restarts = 0
        self.windows_tracked = 0

    def handle_stereo_event(self, event_left: EventArray, event_right: EventArray, msg_timestamp: float):
        """Returns the FeatureCloud published for this pair, or None."""
        if len(event_left) == 0:
            return None                                   # "not event ..." (:150-153)
        if self.first_image_flag:                         # :155-161
            self.first_image_flag = False
            self.first_image_time = msg_timestamp
            self.last_image_time = msg_timestamp
            return None
        if msg_timestamp - self.last_image_time > 1.0 or msg_timestamp < self.last_image_time:
            self.first_image_flag = True                  # :163-173, restart flag published
            self.last_image_time = 0.0
            self.pub_count = 1
            self.restarts += 1
            return None
        self.last_image_time = msg_timestamp
        span = msg_timestamp - self.first_image_time      # frequency control :177-188
        rate = 1.0 * self.pub_count / span if span != 0.0 else math.inf
        if round_half_away(rate) <= self.FREQ:
            pub = True
            if abs(rate - self.FREQ) < 0.01 * self.FREQ:
                self.first_image_time = msg_timestamp
                self.pub_count = 0
        else:
            pub = False
        self.t.PUB_THIS_FRAME = pub
        t_last = float(event_left.t[-1])                  # :190
        if not self.do_mc:
            self.t.trackEvent(t_last, event_left.events, event_right.events)          # :193
        else:
            m = self.motion.assemble(float(event_left.t[0]), event_left.stamp)       # :195-252
            # :254 passes msg_timestamp_left (= t_last, :190); the header stamp is only m["t1"]
            self.t.trackEvent(t_last, event_left.events, event_right.events, m)
        self.windows_tracked += 1
        if not pub:
            return None
        self.pub_count += 1                               # :270
        cloud = FeatureCloud(msg_timestamp, pack_feature_cloud(self.t))
        if not self.init_pub:                             # first publish suppressed (:334-339)
            self.init_pub = True
            return None
        return cloud


class ImagePairer(EventPairer):
    """The depth-1 queues of img_callback_left/right (stereo_image_tracker_node.cpp:36-52) and
    the pairing step of the image node's sync_process (:217-241): +-1 s, and a left frame EXACTLY
    one second older than the right one is already thrown (`<=`, unlike the event node's `<`)."""

    def __init__(self, tolerance: float = 1.0):
        super().__init__(tolerance)

    def poll(self):
        if not self.left or not self.right:
            return None
        tl, tr = self.left[0].stamp, self.right[0].stamp
        if tl <= tr - self.tol:
            self.left.popleft()
            return None
        if tl > tr + self.tol:
            self.right.popleft()
            return None
        l, r = self.left.popleft(), self.right.popleft()
        return l, r, l.stamp


class ImageMsg:
    """header.stamp + the CV_8UC1 frame of one sensor_msgs/Image (after getImageFromMsg)."""

    def __init__(self, stamp: float, image):
        self.stamp = float(stamp)
        self.image = image


class StereoImageNode(StereoEventNode):
    """handle_stereo_image (stereo_image_tracker_node.cpp:54-183) around any tracker with the
    reference's names: the same first-frame skip, restart rule, publish-rate gate, cloud
    packing and first-publish suppression as the event node; the tracker call is
    trackImage(msg_timestamp, img_left, img_right) (:99).  The node's own CLAHE (`EQUALIZE`,
    :93-97) belongs to the tracker behind this class: the GPU frame path applies it to the
    uploaded frames when cfg.equalize is set (esvio_fe_track_image)."""

    def __init__(self, tracker, freq: int):
        super().__init__(tracker, freq)

    def handle_stereo_image(self, img_left, img_right, msg_timestamp: float):
        """Returns the FeatureCloud published for this pair, or None."""
        if self.first_image_flag:                         # :58-64
            self.first_image_flag = False
            self.first_image_time = msg_timestamp
            self.last_image_time = msg_timestamp
            return None
        if msg_timestamp - self.last_image_time > 1.0 or msg_timestamp < self.last_image_time:
            self.first_image_flag = True                  # :66-76
            self.last_image_time = 0.0
            self.pub_count = 1
            self.restarts += 1
            return None
        self.last_image_time = msg_timestamp
        span = msg_timestamp - self.first_image_time      # frequency control :80-91
        rate = 1.0 * self.pub_count / span if span != 0.0 else math.inf
        if round_half_away(rate) <= self.FREQ:
            pub = True
            if abs(rate - self.FREQ) < 0.01 * self.FREQ:
                self.first_image_time = msg_timestamp
                self.pub_count = 0
        else:
            pub = False
        self.t.PUB_THIS_FRAME = pub
        self.t.trackImage(msg_timestamp, img_left, img_right)      # :99
        self.windows_tracked += 1
        if not pub:
            return None
        self.pub_count += 1                               # :113
        cloud = FeatureCloud(msg_timestamp, pack_feature_cloud(self.t))
        if not self.init_pub:                             # :172-177
            self.init_pub = True
            return None
        return cloud


def replay_images(node: StereoImageNode, left_msgs, right_msgs):
    """Plays two ImageMsg lists through the image node's queues and pairing step in stamp
    order (ties: left first), polling after every arrival.  Returns (clouds, overwritten)."""
    pairer = ImagePairer()
    clouds = []
    order = sorted([(m.stamp, 0, i) for i, m in enumerate(left_msgs)]
                   + [(m.stamp, 1, i) for i, m in enumerate(right_msgs)])
    for _, side, i in order:
        if side == 0:
            pairer.push_left(left_msgs[i])
        else:
            pairer.push_right(right_msgs[i])
        while pairer.left and pairer.right:
            pair = pairer.poll()
            if pair is None:
                continue
            c = node.handle_stereo_image(pair[0].image, pair[1].image, pair[2])
            if c is not None:
                clouds.append(c)
    return clouds, pairer.dropped


def round_half_away(v: float) -> float:
    """C round(): half away from zero (Python's round() is half-to-even)."""
    if math.isinf(v) or math.isnan(v):
        return v
    return math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5)


def pack_feature_cloud(t) -> np.ndarray:
    """Rows (x, y, z=1, id*2+cam, u, v, vx, vy) of the `feature` PointCloud (node.cpp:289-329):
    left rows with track_cnt > 1, then right rows whose id was published on the left."""
    ids = np.asarray(t.ids)
    keep = np.asarray(t.track_cnt) > 1
    left = np.zeros((int(keep.sum()), 8), np.float32)
    if len(left):
        left[:, 0:2] = np.asarray(t.cur_un_pts)[keep]
        left[:, 2] = 1.0
        left[:, 3] = (ids[keep] * 2 + 0).astype(np.float32)
        left[:, 4:6] = np.asarray(t.cur_pts)[keep]
        left[:, 6:8] = np.asarray(t.pts_velocity)[keep]
    idr = np.asarray(t.ids_right)
    sel = np.isin(idr, ids[keep]) if len(idr) else np.zeros(0, bool)
    right = np.zeros((int(sel.sum()), 8), np.float32)
    if len(right):
        right[:, 0:2] = np.asarray(t.cur_un_right_pts)[sel]
        right[:, 2] = 1.0
        right[:, 3] = (idr[sel] * 2 + 1).astype(np.float32)
        right[:, 4:6] = np.asarray(t.cur_right_pts)[sel]
        right[:, 6:8] = np.asarray(t.right_pts_velocity)[sel]
    return np.concatenate([left, right], 0)


def decode_feature_cloud(rows: np.ndarray):
    """The consumer's decode (esvio_estimator/src/stereo_estimator_node.cpp:388-401):
    v = ch0 + 0.5; feature_id = v / 2; camera_id = v % 2; z must be 1."""
    v = (rows[:, 3] + 0.5).astype(np.int64)
    assert (rows[:, 2] == 1.0).all()
    return v // 2, v % 2


def window_stream(stream, frequency: float = 30.0, chunk: int = 1 << 20):
    """All EventArray messages EventMessageEditor would write for one raw SoA stream
    (x, y, t, p); the still-open last window is not flushed, as in the tool."""
    w = EventWindower(frequency)
    msgs = []
    n = len(stream[2])
    for i in range(0, n, chunk):
        msgs += w.insert(*(a[i:i + chunk] for a in stream))
    return msgs


def replay(node: StereoEventNode, left_msgs, right_msgs):
    """Plays two message lists through the depth-1 queues, the pairing step and the node in
    header-stamp order (ties: left first), polling after every arrival -- i.e. a consumer that
    always keeps up.  Returns (published clouds, messages overwritten in the queues)."""
    pairer = EventPairer()
    clouds = []
    order = sorted([(m.stamp, 0, i) for i, m in enumerate(left_msgs)]
                   + [(m.stamp, 1, i) for i, m in enumerate(right_msgs)])
    for _, side, i in order:
        if side == 0:
            pairer.push_left(left_msgs[i])
        else:
            pairer.push_right(right_msgs[i])
        while pairer.left and pairer.right:
            pair = pairer.poll()
            if pair is None:
                continue
            c = node.handle_stereo_event(*pair)
            if c is not None:
                clouds.append(c)
    return clouds, pairer.dropped
