"""Deterministic synthetic stereo event streams.

Counter-based splitmix64 so that any window of any stream can be generated
independently and reproducibly.  The scene is a set of axis-aligned rectangles
that translate with constant velocity and wrap at the sensor border; 90 % of the
events sit on rectangle edges (polarity + on leading edges, - on trailing
ones), 10 % are uniform noise.  The right camera sees the same scene shifted by
a per-rectangle disparity towards -x, with independently sampled events.

Timestamps are whole microseconds after T0 = 1.7e9 s, represented the way
`ros::Time::toSec()` produces them (`sec + 1e-9 * nsec`,
/root/reference/feature_tracker/src/feature_tracker.cpp:357 via dvs_msgs/Event),
so float32 time visibly fails and the AoS (sec, nsec) and SoA (f64) forms of the
same event agree bit for bit.
"""
from __future__ import annotations

import numpy as np

T0_SEC = 1_700_000_000
WINDOWS_PER_SEC = 30
_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(seed: int, idx: np.ndarray) -> np.ndarray:
    """z = mix(seed + (idx+1)*golden) for every idx (uint64 in, uint64 out)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + (idx.astype(np.uint64) + np.uint64(1)) * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def _u01(bits: np.ndarray) -> np.ndarray:
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


class Scene:
    """64 rectangles: side 12-40 px, velocity in [-150, 150] px/s, disparity 2-20 px.

    `rigid=True` is the second scene of the benchmarks: the same rectangles as a STATIC world
    seen by a camera that translates parallel to its image plane, so every rectangle moves in
    one common direction with a speed proportional to its disparity (inverse depth).  All
    correspondences then share one epipolar geometry, as on a real sequence, and
    cv::findFundamentalMat (feature_tracker.cpp:935) exits early instead of running its full
    iteration budget the way it does on 64 independently moving objects."""

    def __init__(self, width: int, height: int, seed: int = 42, n_rect: int = 64,
                 max_speed: float = 150.0, rigid: bool = False):
        self.W, self.H, self.n = width, height, n_rect
        r = _u01(splitmix64(seed, np.arange(n_rect * 7))).reshape(n_rect, 7)
        self.w = 12.0 + 28.0 * r[:, 0]
        self.h = 12.0 + 28.0 * r[:, 1]
        self.x0 = r[:, 2] * width
        self.y0 = r[:, 3] * height
        self.vx = (2.0 * r[:, 4] - 1.0) * max_speed
        self.vy = (2.0 * r[:, 5] - 1.0) * max_speed
        self.disp = 2.0 + 18.0 * r[:, 6]
        self.rigid = bool(rigid)
        if rigid:
            theta = 2.0 * np.pi * float(_u01(splitmix64(seed ^ 0x5EED, np.arange(1)))[0])
            speed = max_speed * self.disp / 20.0
            self.vx = speed * np.cos(theta)
            self.vy = speed * np.sin(theta)


class StereoEventStream:
    """One stereo pair at `rate` events/s per camera, cut into 30 windows per second."""

    def __init__(self, width: int, height: int, rate: float, stream: int = 0,
                 scene_seed: int = 42, noise: float = 0.1, mono: bool = False,
                 max_speed: float = 150.0, rigid: bool = False):
        self.W, self.H = width, height
        self.rate = float(rate)
        self.mono = mono
        self.noise = noise
        self.scene = Scene(width, height, scene_seed + 10 * stream, max_speed=max_speed, rigid=rigid)
        self.seeds = (1001 + 10 * stream, 2002 + 10 * stream)
        self.events_per_window = int(round(self.rate / WINDOWS_PER_SEC))

    def window(self, k: int, cam: int, n: int | None = None):
        """Events of window k for camera cam: (x u16, y u16, t f64, p u8, sec u32, nsec u32)."""
        if cam == 1 and self.mono:
            e = np.zeros(0)
            return (e.astype(np.uint16), e.astype(np.uint16), e.astype(np.float64),
                    e.astype(np.uint8), e.astype(np.uint32), e.astype(np.uint32))
        n = self.events_per_window if n is None else n
        sc = self.scene
        seed = self.seeds[cam] * 1_000_003 + k
        bits = splitmix64(seed, np.arange(n * 6)).reshape(n, 6)
        win_us = 1_000_000 // WINDOWS_PER_SEC  # 33 333 us
        us = np.sort((bits[:, 0] % np.uint64(win_us)).astype(np.int64)) + k * win_us
        sec = (T0_SEC + us // 1_000_000).astype(np.uint32)
        nsec = ((us % 1_000_000) * 1000).astype(np.uint32)
        t = sec.astype(np.float64) + 1e-9 * nsec.astype(np.float64)
        tau = us.astype(np.float64) * 1e-6  # seconds since T0

        is_noise = _u01(bits[:, 1]) < self.noise
        rect = (bits[:, 2] % np.uint64(sc.n)).astype(np.int64)
        edge = ((bits[:, 2] >> np.uint64(32)) % np.uint64(4)).astype(np.int64)
        s = _u01(bits[:, 3])
        px = sc.x0[rect] + sc.vx[rect] * tau - (sc.disp[rect] if cam == 1 else 0.0)
        py = sc.y0[rect] + sc.vy[rect] * tau
        w, h = sc.w[rect], sc.h[rect]
        ex = np.where(edge < 2, px + s * w, np.where(edge == 2, px, px + w))
        ey = np.where(edge >= 2, py + s * h, np.where(edge == 0, py, py + h))
        # polarity: + on leading edges (outward normal . velocity > 0)
        vx, vy = sc.vx[rect], sc.vy[rect]
        lead = np.where(edge == 0, vy < 0, np.where(edge == 1, vy > 0,
                        np.where(edge == 2, vx < 0, vx > 0)))
        xi = np.mod(np.rint(ex), self.W).astype(np.int64)
        yi = np.mod(np.rint(ey), self.H).astype(np.int64)
        pol = lead.astype(np.uint8)
        # noise events
        nx = (bits[:, 4] % np.uint64(self.W)).astype(np.int64)
        ny = ((bits[:, 4] >> np.uint64(32)) % np.uint64(self.H)).astype(np.int64)
        npol = (bits[:, 5] & np.uint64(1)).astype(np.uint8)
        x = np.clip(np.where(is_noise, nx, xi), 0, self.W - 1).astype(np.uint16)
        y = np.clip(np.where(is_noise, ny, yi), 0, self.H - 1).astype(np.uint16)
        p = np.where(is_noise, npol, pol).astype(np.uint8)
        return x, y, t, p, sec, nsec

    def stereo_window(self, k: int):
        """((xl, yl, tl, pl), (xr, yr, tr, pr), cur_time) -- cur_time is the last left
        event's time, as at stereo_event_tracker_node.cpp:190."""
        L = self.window(k, 0)
        R = self.window(k, 1)
        cur_time = float(L[2][-1]) if len(L[2]) else 0.0
        return L[:4], R[:4], cur_time


def to_aos(x, y, sec, nsec, p) -> np.ndarray:
    """Pack events the way a std::vector<dvs_msgs::Event> lies in memory (16 B per event:
    u16 x, u16 y, u32 sec, u32 nsec, u8 polarity, 3 B padding;
    /root/reference/feature_tracker/src/dvs_msgs/Event.h:42-52)."""
    dt = np.dtype([("x", "<u2"), ("y", "<u2"), ("sec", "<u4"), ("nsec", "<u4"), ("p", "u1"),
                   ("pad", "V3")])
    assert dt.itemsize == 16
    out = np.zeros(len(x), dt)
    out["x"], out["y"], out["sec"], out["nsec"], out["p"] = x, y, sec, nsec, p
    return out


CAM_DAVIS346 = (
    dict(fx=249.69341447817564, fy=248.41625664694038, cx=176.74240257052816, cy=129.47631010746218,
         k1=-0.3794794654640921, k2=0.15393049046270296, p1=0.0011400586965363895,
         p2=-0.0019042695753031854),
    dict(fx=258.61441518089174, fy=258.00363445501824, cx=178.44356547141308, cy=135.84792628403616,
         k1=-0.3864639588089853, k2=0.1707517912637013, p1=-0.00046695742172563157,
         p2=0.0006610867041757214),
)
CAM_DSEC = (
    dict(fx=553.4686750102932, fy=553.3994078799127, cx=346.65339162053317, cy=216.52092103243012,
         k1=-0.09356476362537607, k2=0.19445779814646236, p1=7.642434980998821e-05,
         p2=0.0019563864604273664),
    dict(fx=552.1819422959984, fy=551.4454720096484, cx=336.87432177064744, cy=226.32630571403274,
         k1=-0.026300, k2=0.037995, p1=-0.000513, p2=0.000167),
)


def default_config(width: int, height: int, **kw) -> dict:
    """Front-end parameter block common to every shipped config."""
    cam = CAM_DAVIS346 if width == 346 else CAM_DSEC
    cfg = dict(width=width, height=height, max_cnt=150, min_dist=10, flow_back=1, equalize=0,
               f_threshold=1.0, ts_lk_threshold=128.0, decay_ms=20.0, ignore_polarity=0,
               median_blur_kernel_size=0, feature_filter_threshold=0.01, focal_length=460.0,
               cam=cam)
    cfg.update(kw)
    return cfg


SCENES = ("survey", "rigid")   # SURVEY.md 8d's 64 independent movers | one rigid world (Scene)

# BASELINE.json configs
WORKLOADS = {
    "mono_davis346_100k": dict(width=346, height=260, rate=1.0e6, mono=True, max_cnt=150, min_dist=10, freq=15),
    "stereo_davis346_1mevs": dict(width=346, height=260, rate=1.0e6, mono=False, max_cnt=150, min_dist=10, freq=15),
    "stereo_vga_5mevs": dict(width=640, height=480, rate=5.0e6, mono=False, max_cnt=150, min_dist=10, freq=10),
    "stereo_vga_20mevs_burst": dict(width=640, height=480, rate=20.0e6, mono=False, max_cnt=200, min_dist=10, freq=10),
    "stereo_vga_10mevs": dict(width=640, height=480, rate=10.0e6, mono=False, max_cnt=150, min_dist=10, freq=10),
}


# ---------------------------------------------------------------------------------------
# frame path (FeatureTracker::trackImage): synthetic intensity images
# ---------------------------------------------------------------------------------------
def frame_texture(W, H, seed):
    """Smooth random texture with corners, built with integer arithmetic only: random blocks
    plus a 5x5 box blur by integer cumulative sums."""
    rng = np.random.default_rng(seed)
    coarse = rng.integers(0, 256, ((H + 15) // 8 + 2, (W + 15) // 8 + 2), dtype=np.int64)
    img = np.kron(coarse, np.ones((8, 8), np.int64))[:H + 4, :W + 4]
    img = img + rng.integers(0, 24, img.shape, dtype=np.int64)
    c = np.cumsum(np.cumsum(np.pad(img, ((1, 0), (1, 0))), axis=0), axis=1)
    box = c[5:, 5:] - c[:-5, 5:] - c[5:, :-5] + c[:-5, :-5]
    return ((box + 12) // 25).clip(0, 255).astype(np.uint8)[:H, :W]


def stereo_frame_sequence(W, H, n_frames, seed=5):
    """Frames cropped from one big texture at integer offsets: camera pans (3, 2) px per
    frame, the right camera sees the scene 6 px further left."""
    big = frame_texture(W + 8 * n_frames + 32, H + 8 * n_frames + 32, seed)
    out = []
    for k in range(n_frames):
        ox, oy = 8 + 3 * k, 8 + 2 * k
        out.append((np.ascontiguousarray(big[oy:oy + H, ox:ox + W]),
                    np.ascontiguousarray(big[oy:oy + H, ox + 6:ox + 6 + W])))
    return out
