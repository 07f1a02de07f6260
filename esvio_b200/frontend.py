"""Host-side mirror of the reference's front-end interface on top of the C ABI.

`FeatureTracker` keeps the names of the reference class
(/root/reference/feature_tracker/src/feature_tracker.h:44-135): `trackEvent(cur_time,
event_left, event_right)` followed by the public result members the node reads
(`ids, track_cnt, cur_pts, cur_un_pts, pts_velocity, ids_right, cur_right_pts,
cur_un_right_pts, right_pts_velocity`; stereo_event_tracker_node.cpp:289-323).  All work
happens in libesvio_fe.so on the GPU; this file only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import Config, Events, FrontEndError, Motion, Pinhole, Tracks

AOS_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("sec", "<u4"), ("nsec", "<u4"), ("p", "u1"),
                      ("pad", "V3")])


def pipeline_depth() -> int:
    """Windows that may be in flight between submit() and wait() (esvio_fe_pipeline_depth)."""
    return int(_capi.lib().esvio_fe_pipeline_depth())


def nccl_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0; ship the 128 bytes to the other ranks)."""
    buf = C.create_string_buffer(128)
    st = _capi.lib().esvio_fe_nccl_unique_id(buf)
    if st != _capi.OK:
        raise FrontEndError(st, "nccl_unique_id")
    return buf.raw


def make_config(cfg: dict) -> Config:
    """dict with the reference's parameter names -> esvio_fe_config."""
    c = Config()
    _capi.lib().esvio_fe_default_config(C.byref(c), int(cfg["width"]), int(cfg["height"]))
    for k in ("max_cnt", "min_dist", "flow_back", "equalize", "ignore_polarity",
              "median_blur_kernel_size", "do_motion_correction", "device_id",
              "max_events_per_window", "use_ransac"):
        if k in cfg:
            setattr(c, k, int(cfg[k]))
    for k in ("f_threshold", "ts_lk_threshold", "decay_ms", "feature_filter_threshold",
              "focal_length", "mc_fx", "mc_fy", "mc_cx", "mc_cy"):
        if k in cfg:
            setattr(c, k, float(cfg[k]))
    if "cam" in cfg:
        for i in range(2):
            cam = cfg["cam"][i]
            c.cam[i] = Pinhole(*[float(cam[k]) for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")])
    return c


def make_motion(m) -> Motion:
    """dict(state_v, v_pre, accel, omega, t1) -> esvio_motion (Motion_correction_value,
    feature_tracker.h:35)."""
    if isinstance(m, Motion):
        return m
    o = Motion()
    for k in ("state_v", "v_pre", "accel", "omega"):
        for i in range(3):
            getattr(o, k)[i] = float(m[k][i])
    o.t1 = float(m["t1"])
    return o


class _Ev:
    """Keeps the numpy arrays alive next to the esvio_events struct that points into them."""

    def __init__(self, ev):
        self.s = Events()
        self.keep = None
        if ev is None:
            return
        if isinstance(ev, DeviceEvents):
            self.s.x, self.s.y, self.s.t, self.s.p = ev.x, ev.y, ev.t, ev.p
            self.s.aos = ev.aos
            self.s.n = ev.n
            self.s.on_device = 1
            return
        if isinstance(ev, np.ndarray) and ev.dtype == AOS_DTYPE:
            a = np.ascontiguousarray(ev)
            self.keep = (a,)
            self.s.aos = a.ctypes.data
            self.s.n = len(a)
            return
        if getattr(ev, "stereo_block", False):
            self.s.flags = _capi.EVENTS_STEREO_BLOCK
        x, y, t, p = ev[:4]
        x = np.ascontiguousarray(x, np.uint16)
        y = np.ascontiguousarray(y, np.uint16)
        t = np.ascontiguousarray(t, np.float64)
        p = np.ascontiguousarray(p, np.uint8)
        self.keep = (x, y, t, p)
        self.s.x, self.s.y, self.s.t, self.s.p = (a.ctypes.data for a in self.keep)
        self.s.n = len(x)


class DeviceEvents:
    """One camera's events of one window, resident in HBM (esvio_events.on_device = 1)."""

    def __init__(self, fe: "EventFrontEnd", ev):
        self.fe = fe
        self.x = self.y = self.t = self.p = self.aos = None
        self._ptrs = []
        if isinstance(ev, np.ndarray) and ev.dtype == AOS_DTYPE:
            self.n = len(ev)
            self.aos = self._up(np.ascontiguousarray(ev))
        else:
            x, y, t, p = ev[:4]
            self.n = len(x)
            self.x = self._up(np.ascontiguousarray(x, np.uint16))
            self.y = self._up(np.ascontiguousarray(y, np.uint16))
            self.t = self._up(np.ascontiguousarray(t, np.float64))
            self.p = self._up(np.ascontiguousarray(p, np.uint8))

    def _up(self, a):
        ptr = C.c_void_p()
        self.fe._chk(_capi.lib().esvio_fe_device_alloc(self.fe._h, max(a.nbytes, 1), C.byref(ptr)),
                     "device_alloc")
        if a.nbytes:
            self.fe._chk(_capi.lib().esvio_fe_copy_to_device(self.fe._h, ptr, a.ctypes.data,
                                                             a.nbytes), "copy_to_device")
        self._ptrs.append(ptr)
        return ptr.value

    def free(self):
        for p in self._ptrs:
            _capi.lib().esvio_fe_device_free(self.fe._h, p)
        self._ptrs = []


class PinnedEvents:
    """SoA event arrays in ONE block of pinned host memory (esvio_fe_host_alloc) laid out by
    esvio_fe_soa_layout, viewed as numpy: the window crosses PCIe as a single copy."""

    def __init__(self, ev):
        x, y, t, p = ev[:4]
        n = len(x)
        self.n = n
        L = _capi.lib()
        off = (C.c_size_t * 4)()
        total = C.c_size_t()
        L.esvio_fe_soa_layout(n, off, C.byref(total))
        ptr = L.esvio_fe_host_alloc(max(total.value, 16))
        if not ptr:
            raise MemoryError("esvio_fe_host_alloc failed")
        self._raw = [ptr]
        self._buf = (C.c_uint8 * max(total.value, 16)).from_address(ptr)
        views = []
        for a, dt, o in zip((x, y, t, p), (np.uint16, np.uint16, np.float64, np.uint8), off):
            v = np.frombuffer(self._buf, dtype=dt, count=n, offset=o)
            v[:] = np.ascontiguousarray(a, dt)
            views.append(v)
        self.arrays = tuple(views)

    def __getitem__(self, i):
        return self.arrays[i]

    def free(self):
        for p in self._raw:
            _capi.lib().esvio_fe_host_free(p)
        self._raw = []


class _PinnedView:
    def __init__(self, arrays):
        self.arrays = tuple(arrays)
        self.n = len(arrays[0])

    def __getitem__(self, i):
        return self.arrays[i]


class PinnedStereoEvents:
    """Both cameras' SoA event arrays of a window in ONE block of pinned host memory laid out by
    esvio_fe_soa_layout_stereo: the window crosses PCIe as a single copy.  `.left` / `.right`
    are (x, y, t, p) views to hand to submit()."""

    def __init__(self, left, right):
        L = _capi.lib()
        nl, nr = len(left[0]), len(right[0])
        ol, orr = (C.c_size_t * 4)(), (C.c_size_t * 4)()
        total = C.c_size_t()
        L.esvio_fe_soa_layout_stereo(nl, nr, ol, orr, C.byref(total))
        ptr = L.esvio_fe_host_alloc(max(total.value, 16))
        if not ptr:
            raise MemoryError("esvio_fe_host_alloc failed")
        self._raw = [ptr]
        self._buf = (C.c_uint8 * max(total.value, 16)).from_address(ptr)
        sides = []
        for ev, n, off in ((left, nl, ol), (right, nr, orr)):
            views = []
            for a, dt, o in zip(ev[:4], (np.uint16, np.uint16, np.float64, np.uint8), off):
                v = np.frombuffer(self._buf, dtype=dt, count=n, offset=o)
                v[:] = np.ascontiguousarray(a, dt)
                views.append(v)
            sides.append(_PinnedView(views))
        self.left, self.right = sides
        self.left.stereo_block = self.right.stereo_block = True

    def free(self):
        for p in self._raw:
            _capi.lib().esvio_fe_host_free(p)
        self._raw = []


class EventFrontEnd:
    """Thin object wrapper around one `esvio_fe` handle (one stereo event stream)."""

    def __init__(self, cfg: dict):
        self.cfg = dict(cfg)
        self._c = make_config(cfg)
        self.W, self.H, self.M = self._c.width, self._c.height, self._c.max_cnt
        h = C.c_void_p()
        st = _capi.lib().esvio_fe_create(C.byref(self._c), C.byref(h))
        if st != _capi.OK:
            raise FrontEndError(st, "esvio_fe_create")
        self._h = h
        M = self.M
        self._bufs = {k: np.zeros(M, np.int32) for k in ("id", "track_cnt", "id_right")}
        self._bufs.update({k: np.zeros(M, np.float32) for k in
                           ("u", "v", "un_x", "un_y", "vx", "vy", "ru", "rv", "run_x", "run_y",
                            "rvx", "rvy")})
        self._t = Tracks()
        self._t.capacity = M
        for k, a in self._bufs.items():
            setattr(self._t, k, a.ctypes.data_as(_capi._pi if a.dtype == np.int32 else _capi._pf))
        self._inflight = []

    def close(self):
        if getattr(self, "_h", None):
            _capi.lib().esvio_fe_destroy(self._h)
            self._h = None

    def __del__(self):
        if _capi is not None:   # at interpreter shutdown the module globals may already be gone
            self.close()

    def _chk(self, st, where):
        if st != _capi.OK:
            raise FrontEndError(st, where, _capi.lib().esvio_fe_last_error(self._h).decode())

    # ---- hot path
    def _unpack(self):
        t, b = self._t, self._bufs
        nl, nr = t.n_left, t.n_right
        out = {k: b[k][:nl].copy() for k in ("id", "track_cnt", "u", "v", "un_x", "un_y", "vx", "vy")}
        out.update({k: b[k][:nr].copy() for k in ("id_right", "ru", "rv", "run_x", "run_y", "rvx", "rvy")})
        s = t.stats
        out["stats"] = dict(n_prev=s.n_prev, n_after_temporal=s.n_after_temporal,
                            n_after_ransac=s.n_after_ransac, n_after_mask=s.n_after_mask,
                            n_new=s.n_new, ransac_iters=s.ransac_iters)
        return out

    def track(self, cur_time, left, right, pub_this_frame=True, motion=None):
        l, r = _Ev(left), _Ev(right)
        m = make_motion(motion) if motion is not None else None
        self._chk(_capi.lib().esvio_fe_track_mc(self._h, float(cur_time), C.byref(l.s), C.byref(r.s),
                                                int(bool(pub_this_frame)),
                                                C.byref(m) if m is not None else None,
                                                C.byref(self._t)), "track")
        return self._unpack()

    def track_raw(self, cur_time, l: _Ev, r: _Ev, pub_this_frame=True):
        """Same call with pre-built esvio_events structs; returns (n_left, n_right)."""
        self._chk(_capi.lib().esvio_fe_track(self._h, float(cur_time), C.byref(l.s), C.byref(r.s),
                                             int(bool(pub_this_frame)), C.byref(self._t)), "track")
        return self._t.n_left, self._t.n_right

    def submit(self, cur_time, left, right, pub_this_frame=True):
        l = left if isinstance(left, _Ev) else _Ev(left)
        r = right if isinstance(right, _Ev) else _Ev(right)
        self._inflight.append((l, r))
        self._chk(_capi.lib().esvio_fe_track_submit(self._h, float(cur_time), C.byref(l.s),
                                                    C.byref(r.s), int(bool(pub_this_frame))),
                  "track_submit")

    def wait(self, unpack=True):
        self._chk(_capi.lib().esvio_fe_track_wait(self._h, C.byref(self._t)), "track_wait")
        self._inflight.pop(0)
        return self._unpack() if unpack else (self._t.n_left, self._t.n_right)

    # ---- left/right split over two GPUs (include/esvio_fe.h "Left/right split")
    def split_image_submit(self, cur_time, events, exchange_stream=0):
        """Right GPU: SAE update + time surface + pyramid of this handle's one camera; returns
        (device pointer, bytes) of the image block, ready on `exchange_stream` (cudaStream_t)."""
        e = events if isinstance(events, _Ev) else _Ev(events)
        self._split_keep = getattr(self, "_split_keep", [])[-(pipeline_depth() - 1):] + [e]   # host buffers of the windows in flight
        p, n = C.c_void_p(), C.c_size_t()
        self._chk(_capi.lib().esvio_fe_split_image_submit(self._h, float(cur_time), C.byref(e.s),
                                                          C.c_void_p(exchange_stream), C.byref(p),
                                                          C.byref(n)), "split_image_submit")
        return p.value, n.value

    def split_right_buffer(self):
        """Left GPU: (device pointer, bytes) the next window's right image must be written to."""
        p, n = C.c_void_p(), C.c_size_t()
        self._chk(_capi.lib().esvio_fe_split_right_buffer(self._h, C.byref(p), C.byref(n)),
                  "split_right_buffer")
        return p.value, n.value

    def submit_split(self, cur_time, left, pub_this_frame=True, exchange_stream=0):
        """Left GPU: track_submit with the right image taken from split_right_buffer()."""
        l = left if isinstance(left, _Ev) else _Ev(left)
        self._inflight.append((l, None))
        self._chk(_capi.lib().esvio_fe_track_submit_split(self._h, float(cur_time), C.byref(l.s),
                                                          int(bool(pub_this_frame)),
                                                          C.c_void_p(exchange_stream)),
                  "track_submit_split")

    # ---- frame path (FeatureTracker::trackImage)
    def _img(self, a):
        a = np.ascontiguousarray(a, np.uint8)
        if a.shape != (self.H, self.W):
            raise ValueError(f"image must be {self.H}x{self.W} uint8")
        return a

    def track_image(self, cur_time, img_left, img_right=None, pub_this_frame=True):
        l = self._img(img_left)
        r = self._img(img_right) if img_right is not None else None
        self._chk(_capi.lib().esvio_fe_track_image(
            self._h, float(cur_time), l.ctypes.data, self.W, r.ctypes.data if r is not None else None,
            self.W, int(bool(pub_this_frame)), C.byref(self._t)), "track_image")
        return self._unpack()

    def submit_image(self, cur_time, img_left, img_right=None, pub_this_frame=True):
        l = self._img(img_left)
        r = self._img(img_right) if img_right is not None else None
        self._inflight.append((l, r))       # the host frames stay alive until the wait
        self._chk(_capi.lib().esvio_fe_track_image_submit(
            self._h, float(cur_time), l.ctypes.data, self.W, r.ctypes.data if r is not None else None,
            self.W, int(bool(pub_this_frame))), "track_image_submit")

    def stage_good_features(self, img, max_corners, min_distance, mask=None, want_eig=False):
        a = self._img(img)
        m = self._img(mask) if mask is not None else None
        cap = max_corners if max_corners > 0 else self.W * self.H
        out = np.zeros((max(cap, 1), 2), np.float32)
        n = C.c_int32()
        eig = np.empty((self.H, self.W), np.float32) if want_eig else None
        self._chk(_capi.lib().esvio_fe_stage_good_features(
            self._h, a.ctypes.data, m.ctypes.data if m is not None else None, int(max_corners),
            float(min_distance), out.ctypes.data, cap, C.byref(n),
            eig.ctypes.data if eig is not None else None), "stage_good_features")
        pts = out[:min(n.value, cap)].copy()
        return (pts, eig) if want_eig else pts

    def reset(self):
        self._chk(_capi.lib().esvio_fe_reset(self._h), "reset")

    def time_surface(self, cam):
        out = np.empty((self.H, self.W), np.uint8)
        self._chk(_capi.lib().esvio_fe_time_surface(self._h, cam, out.ctypes.data, self.W),
                  "time_surface")
        return out

    # ---- profiling / plumbing
    def set_profiling(self, on=True):
        self._chk(_capi.lib().esvio_fe_set_profiling(self._h, int(on)), "set_profiling")

    def stage_ms(self):
        ms = (C.c_float * _capi.NUM_STAGES)()
        self._chk(_capi.lib().esvio_fe_get_stage_ms(self._h, ms), "get_stage_ms")
        return dict(zip(_capi.STAGE_NAMES, list(ms)))

    def stage_marks(self):
        """Timeline of the last profiled window: ms since set_profiling(True) of the 12 markers
        (include/esvio_fe.h)."""
        ms = (C.c_float * 13)()
        self._chk(_capi.lib().esvio_fe_get_stage_marks(self._h, ms), "get_stage_marks")
        return list(ms)

    def kernel_launches(self):
        n = C.c_int64()
        self._chk(_capi.lib().esvio_fe_kernel_launches(self._h, C.byref(n)), "kernel_launches")
        return n.value

    def result_device_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._chk(_capi.lib().esvio_fe_result_device_ptr(self._h, C.byref(p), C.byref(n)),
                  "result_device_ptr")
        return p.value, n.value

    def result_acquire(self, consumer_stream=0):
        """(device pointer, bytes) of the last submitted window's packed records, ordered on
        `consumer_stream` (cudaStream_t); pair with result_release()."""
        p, n = C.c_void_p(), C.c_size_t()
        self._chk(_capi.lib().esvio_fe_result_acquire(self._h, C.c_void_p(consumer_stream),
                                                      C.byref(p), C.byref(n)), "result_acquire")
        return p.value, n.value

    def result_release(self, consumer_stream=0):
        self._chk(_capi.lib().esvio_fe_result_release(self._h, C.c_void_p(consumer_stream)),
                  "result_release")

    # ---- time-window shard building blocks (include/esvio_fe.h, esvio_b200/shard.py)
    def state_device_ptrs(self):
        """(sae pointer, sae_latest pointer, bytes each): double2[2 cams][H][W] planes in HBM."""
        a, b, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
        self._chk(_capi.lib().esvio_fe_state_device_ptrs(self._h, C.byref(a), C.byref(b), C.byref(n)),
                  "state_device_ptrs")
        return a.value, b.value, n.value

    def shard_merge_max(self, dst, srcs, n_doubles, stream=0):
        arr = (C.c_void_p * len(srcs))(*[int(p) for p in srcs])
        self._chk(_capi.lib().esvio_fe_shard_merge_max(self._h, C.c_void_p(int(dst)), arr, len(srcs),
                                                       int(n_doubles), C.c_void_p(stream)), "shard_merge_max")

    def shard_event_stage(self, t_ref, left, right, stream=0):
        l = left if isinstance(left, _Ev) else _Ev(left)
        r = right if isinstance(right, _Ev) else _Ev(right)
        self._shard_keep = (l, r)
        self._chk(_capi.lib().esvio_fe_shard_event_stage(self._h, float(t_ref), C.byref(l.s), C.byref(r.s),
                                                         C.c_void_p(stream)), "shard_event_stage")

    def shard_corner_candidates(self, left, stream=0):
        l = left if isinstance(left, _Ev) else _Ev(left)
        a, b = C.c_void_p(), C.c_void_p()
        self._chk(_capi.lib().esvio_fe_shard_corner_candidates(self._h, C.byref(l.s), C.c_void_p(stream),
                                                               C.byref(a), C.byref(b)), "shard_corner_candidates")
        return a.value, b.value

    def shard_sizes(self):
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        self._chk(_capi.lib().esvio_fe_shard_sizes(self._h, C.byref(a), C.byref(b), C.byref(c)), "shard_sizes")
        return a.value, b.value, c.value

    def shard_images(self):
        a, b = C.c_void_p(), C.c_void_p()
        self._chk(_capi.lib().esvio_fe_shard_images(self._h, C.byref(a), C.byref(b)), "shard_images")
        return a.value, b.value

    def external_buffers(self):
        p = [C.c_void_p() for _ in range(4)]
        self._chk(_capi.lib().esvio_fe_external_buffers(self._h, *[C.byref(x) for x in p]), "external_buffers")
        return tuple(x.value for x in p)

    def submit_external(self, cur_time, n_left_events, pub_this_frame=True, stream=0):
        self._inflight.append((None, None))
        self._chk(_capi.lib().esvio_fe_track_submit_external(self._h, float(cur_time), int(n_left_events),
                                                             int(bool(pub_this_frame)), C.c_void_p(stream)),
                  "track_submit_external")

    # ---- replica mode: all-gather of the packed track records over NCCL (include/esvio_fe.h)
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._chk(_capi.lib().esvio_fe_comm_init(self._h, buf, int(rank), int(world)), "comm_init")

    def allgather_tracks(self):
        self._chk(_capi.lib().esvio_fe_allgather_tracks(self._h), "allgather_tracks")

    def gathered_tracks(self):
        """(device pointer of [world] blocks, bytes per rank, cudaStream_t of the collective)."""
        p, n, st = C.c_void_p(), C.c_size_t(), C.c_void_p()
        self._chk(_capi.lib().esvio_fe_gathered_tracks(self._h, C.byref(p), C.byref(n), C.byref(st)),
                  "gathered_tracks")
        return p.value, n.value, st.value or 0

    def stream(self):
        s = C.c_void_p()
        self._chk(_capi.lib().esvio_fe_stream(self._h, C.byref(s)), "stream")
        return s.value or 0

    # ---- stage-level entry points (parity tests)
    def sae_planes(self, cam):
        """(sae[0], sae[1], latest[0], latest[1]) as HxW float64."""
        out = []
        for plane in range(4):
            a = np.empty((self.H, self.W), np.float64)
            self._chk(_capi.lib().esvio_fe_get_sae(self._h, cam, plane, a.ctypes.data), "get_sae")
            out.append(a)
        return out

    def stage_update(self, t_ref, left, right, motion=None):
        l, r = _Ev(left), _Ev(right)
        m = make_motion(motion) if motion is not None else None
        self._chk(_capi.lib().esvio_fe_stage_update_mc(self._h, float(t_ref), C.byref(l.s),
                                                       C.byref(r.s),
                                                       C.byref(m) if m is not None else None),
                  "stage_update")

    def stage_motion_correct(self, motion, xy_dt):
        """EventDetector::motioncorrection on (x, y, dt) triples -> (n, 2) int pixels."""
        a = np.ascontiguousarray(xy_dt, np.float32).reshape(-1, 3)
        out = np.zeros((len(a), 2), np.int32)
        m = make_motion(motion)
        self._chk(_capi.lib().esvio_fe_stage_motion_correct(self._h, C.byref(m), a.ctypes.data,
                                                            len(a), out.ctypes.data),
                  "stage_motion_correct")
        return out

    def stage_corner_flags(self, left, and_ts_test=False):
        l = _Ev(left)
        out = np.zeros(max(l.s.n, 1), np.uint8)
        self._chk(_capi.lib().esvio_fe_stage_corner_flags(self._h, C.byref(l.s), int(and_ts_test),
                                                          out.ctypes.data), "stage_corner_flags")
        return out[:l.s.n]

    def pyramid_level(self, which, level):
        w, h = C.c_int32(), C.c_int32()
        self._chk(_capi.lib().esvio_fe_get_pyramid_level(self._h, which, level, None, C.byref(w),
                                                         C.byref(h)), "get_pyramid_level")
        out = np.empty((h.value, w.value), np.uint8)
        self._chk(_capi.lib().esvio_fe_get_pyramid_level(self._h, which, level, out.ctypes.data,
                                                         C.byref(w), C.byref(h)),
                  "get_pyramid_level")
        return out

    def stage_lk(self, prev_img, next_img, prev_pts, next_pts=None, max_level=3):
        a = np.ascontiguousarray(prev_img, np.uint8)
        b = np.ascontiguousarray(next_img, np.uint8)
        assert a.shape == (self.H, self.W) and b.shape == (self.H, self.W)
        pp = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
        n = len(pp)
        init = next_pts is not None
        npts = (np.ascontiguousarray(next_pts, np.float32).reshape(-1, 2).copy() if init
                else np.zeros((n, 2), np.float32))
        st = np.zeros(max(n, 1), np.uint8)
        self._chk(_capi.lib().esvio_fe_stage_lk(self._h, a.ctypes.data, b.ctypes.data,
                                                pp.ctypes.data, npts.ctypes.data, n,
                                                st.ctypes.data, max_level, int(init)), "stage_lk")
        return npts, st[:n]

    def stage_condition(self, img, median_ksize=0, equalize=False):
        a = np.ascontiguousarray(img, np.uint8)
        assert a.shape == (self.H, self.W)
        out = np.empty_like(a)
        self._chk(_capi.lib().esvio_fe_stage_condition(self._h, a.ctypes.data, int(median_ksize),
                                                       int(bool(equalize)), out.ctypes.data),
                  "stage_condition")
        return out

    def stage_fmat_mask(self, p1, p2, thresh=1.0):
        p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 2)
        p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 2)
        n = len(p1)
        mask = np.zeros(max(n, 1), np.uint8)
        it = C.c_int32()
        self._chk(_capi.lib().esvio_fe_stage_fmat_mask(self._h, p1.ctypes.data, p2.ctypes.data, n,
                                                       float(thresh), mask.ctypes.data,
                                                       C.byref(it)), "stage_fmat_mask")
        return mask[:n], it.value

    def stage_select(self, left, pts, ids, track_cnt):
        l = _Ev(left)
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        ids = np.ascontiguousarray(ids, np.int32)
        cnt = np.ascontiguousarray(track_cnt, np.int32)
        n = len(ids)
        M = self.M
        po, io, co = np.zeros((M, 2), np.float32), np.zeros(M, np.int32), np.zeros(M, np.int32)
        n_out, n_kept = C.c_int32(), C.c_int32()
        self._chk(_capi.lib().esvio_fe_stage_select(
            self._h, C.byref(l.s), n, pts.ctypes.data, ids.ctypes.data, cnt.ctypes.data,
            C.byref(n_out), po.ctypes.data, io.ctypes.data, co.ctypes.data, C.byref(n_kept)),
            "stage_select")
        k = n_out.value
        return po[:k], io[:k], co[:k], n_kept.value

    def stage_set_tracks(self, prev_time, next_id, tracks):
        """Teacher forcing: the tracker's carried state becomes `tracks` (a result dict of this
        class or of the oracle: id, track_cnt, u, v, un_x, un_y, id_right, run_x, run_y)."""
        n, nr = len(tracks["id"]), len(tracks["id_right"])
        pts = np.ascontiguousarray(np.stack([tracks["u"], tracks["v"]], 1), np.float32)
        un = np.ascontiguousarray(np.stack([tracks["un_x"], tracks["un_y"]], 1), np.float32)
        unr = np.ascontiguousarray(np.stack([tracks["run_x"], tracks["run_y"]], 1), np.float32)
        ids = np.ascontiguousarray(tracks["id"], np.int32)
        cnt = np.ascontiguousarray(tracks["track_cnt"], np.int32)
        idr = np.ascontiguousarray(tracks["id_right"], np.int32)
        self._chk(_capi.lib().esvio_fe_stage_set_tracks(
            self._h, float(prev_time), int(next_id), n, pts.ctypes.data, ids.ctypes.data,
            cnt.ctypes.data, un.ctypes.data, nr, idr.ctypes.data, unr.ctypes.data), "stage_set_tracks")

    def stage_sort_order(self, key, depth_limit=-1):
        """Visiting order of Event_setMask / Image_setMask (libstdc++'s std::sort, key descending)."""
        key = np.ascontiguousarray(key, np.int32)
        out = np.zeros(len(key), np.int32)
        self._chk(_capi.lib().esvio_fe_stage_sort_order(self._h, key.ctypes.data, len(key), int(depth_limit),
                                                        out.ctypes.data), "stage_sort_order")
        return out

    def stage_undistort(self, cam, uv):
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        out = np.zeros_like(uv)
        self._chk(_capi.lib().esvio_fe_stage_undistort(self._h, cam, uv.ctypes.data, len(uv),
                                                       out.ctypes.data), "stage_undistort")
        return out


class _MemberView(EventFrontEnd):
    """Read-only view of a group member (time surface, SAE planes, pyramids)."""

    def __init__(self, handle, cfg_struct):
        self._c = cfg_struct
        self.W, self.H, self.M = cfg_struct.width, cfg_struct.height, cfg_struct.max_cnt
        self._h = C.c_void_p(handle)
        self._inflight = []

    def close(self):
        self._h = None


class EventFrontEndGroup:
    """`esvio_fe_group`: S independent stereo streams of one configuration whose event stage
    runs batched (one k_sae_update_ts launch per window for all 2S cameras)."""

    def __init__(self, cfg: dict, n_streams: int):
        self.S = int(n_streams)
        self._c = make_config(cfg)
        self.M = self._c.max_cnt
        h = C.c_void_p()
        st = _capi.lib().esvio_fe_group_create(C.byref(self._c), self.S, C.byref(h))
        if st != _capi.OK:
            raise FrontEndError(st, "esvio_fe_group_create")
        self._h = h
        self._tracks = (Tracks * self.S)()
        self._bufs = []
        for i in range(self.S):
            b = {k: np.zeros(self.M, np.int32) for k in ("id", "track_cnt", "id_right")}
            b.update({k: np.zeros(self.M, np.float32) for k in
                      ("u", "v", "un_x", "un_y", "vx", "vy", "ru", "rv", "run_x", "run_y", "rvx", "rvy")})
            self._tracks[i].capacity = self.M
            for k, a in b.items():
                setattr(self._tracks[i], k, a.ctypes.data_as(_capi._pi if a.dtype == np.int32 else _capi._pf))
            self._bufs.append(b)
        self._inflight = []

    def close(self):
        if getattr(self, "_h", None):
            _capi.lib().esvio_fe_group_destroy(self._h)
            self._h = None

    def __del__(self):
        if _capi is not None:
            self.close()

    def reset(self):
        self._chk(_capi.lib().esvio_fe_group_reset(self._h), "group_reset")

    def member(self, i) -> EventFrontEnd:
        return _MemberView(_capi.lib().esvio_fe_group_member(self._h, int(i)), self._c)

    def _chk(self, st, where):
        if st != _capi.OK:
            m = _capi.lib().esvio_fe_group_member(self._h, 0)
            raise FrontEndError(st, where, _capi.lib().esvio_fe_last_error(m).decode())

    def submit(self, cur_times, lefts, rights, pubs):
        ls = [l if isinstance(l, _Ev) else _Ev(l) for l in lefts]
        rs = [r if isinstance(r, _Ev) else _Ev(r) for r in rights]
        la, ra = (Events * self.S)(*[e.s for e in ls]), (Events * self.S)(*[e.s for e in rs])
        ct = (C.c_double * self.S)(*[float(t) for t in cur_times])
        pb = (C.c_int32 * self.S)(*[int(bool(p)) for p in pubs])
        self._inflight.append((ls, rs, la, ra))
        self._chk(_capi.lib().esvio_fe_group_track_submit(self._h, ct, la, ra, pb), "group_track_submit")

    def wait(self, unpack=True):
        self._chk(_capi.lib().esvio_fe_group_track_wait(self._h, self._tracks), "group_track_wait")
        self._inflight.pop(0)
        if not unpack:
            return [(t.n_left, t.n_right) for t in self._tracks]
        out = []
        for t, b in zip(self._tracks, self._bufs):
            nl, nr = t.n_left, t.n_right
            o = {k: b[k][:nl].copy() for k in ("id", "track_cnt", "u", "v", "un_x", "un_y", "vx", "vy")}
            o.update({k: b[k][:nr].copy() for k in ("id_right", "ru", "rv", "run_x", "run_y", "rvx", "rvy")})
            out.append(o)
        return out

    def track(self, cur_times, lefts, rights, pubs):
        self.submit(cur_times, lefts, rights, pubs)
        return self.wait()

    def kernel_launches(self):
        n = C.c_int64()
        self._chk(_capi.lib().esvio_fe_group_kernel_launches(self._h, C.byref(n)), "group_kernel_launches")
        return n.value

    def sae_ts_ms(self):
        ms = C.c_float()
        self._chk(_capi.lib().esvio_fe_group_sae_ts_ms(self._h, C.byref(ms)), "group_sae_ts_ms")
        return ms.value


class FeatureTracker:
    """Drop-in mirror of the reference's `FeatureTracker` for the event path
    (feature_tracker.h:44-135).  `PUB_THIS_FRAME` is the reference's global of the same name
    (parameters.h; set by the node at stereo_event_tracker_node.cpp:179,188)."""

    def __init__(self, cfg: dict):
        self.fe = EventFrontEnd(cfg)
        self.PUB_THIS_FRAME = True
        self.cur_time = 0.0
        self.prev_time = 0.0
        self.ids = np.zeros(0, np.int32)
        self.track_cnt = np.zeros(0, np.int32)
        self.cur_pts = np.zeros((0, 2), np.float32)
        self.cur_un_pts = np.zeros((0, 2), np.float32)
        self.pts_velocity = np.zeros((0, 2), np.float32)
        self.ids_right = np.zeros(0, np.int32)
        self.cur_right_pts = np.zeros((0, 2), np.float32)
        self.cur_un_right_pts = np.zeros((0, 2), np.float32)
        self.right_pts_velocity = np.zeros((0, 2), np.float32)
        self.stats = {}

    def trackEvent(self, _cur_time, event_left, event_right, measurements=None):
        """Both overloads of FeatureTracker::trackEvent (feature_tracker.h:51-52):
        `measurements` is the Motion_correction_value of the motion-compensated one."""
        r = self.fe.track(_cur_time, event_left, event_right, self.PUB_THIS_FRAME, measurements)
        self._take(_cur_time, r)

    def trackImage(self, _cur_time, img_left, img_right=None):
        """FeatureTracker::trackImage (feature_tracker.h:49): CV_8UC1 frames of the configured
        size; img_right None (or empty) is the mono case.  Use a tracker of its own for frames."""
        if img_right is not None and np.size(img_right) == 0:
            img_right = None
        r = self.fe.track_image(_cur_time, img_left, img_right, self.PUB_THIS_FRAME)
        self._take(_cur_time, r)

    def _take(self, _cur_time, r):
        self.prev_time, self.cur_time = self.cur_time, float(_cur_time)
        self.ids, self.track_cnt = r["id"], r["track_cnt"]
        self.cur_pts = np.stack([r["u"], r["v"]], 1)
        self.cur_un_pts = np.stack([r["un_x"], r["un_y"]], 1)
        self.pts_velocity = np.stack([r["vx"], r["vy"]], 1)
        self.ids_right = r["id_right"]
        self.cur_right_pts = np.stack([r["ru"], r["rv"]], 1)
        self.cur_un_right_pts = np.stack([r["run_x"], r["run_y"]], 1)
        self.right_pts_velocity = np.stack([r["rvx"], r["rvy"]], 1)
        self.stats = r["stats"]

    def gettimesurface(self):
        """feature_tracker.cpp:894-897"""
        return self.fe.time_surface(0)

    def feature_point_cloud(self):
        """Rows of the `feature` sensor_msgs/PointCloud exactly as the node packs them
        (stereo_event_tracker_node.cpp:268-329): (x, y, z=1, id*2+cam, u, v, vx, vy), left rows
        with track_cnt > 1 first, then right rows whose id was published on the left."""
        rows = []
        pub = set()
        for j in range(len(self.ids)):
            if self.track_cnt[j] > 1:
                pid = int(self.ids[j])
                pub.add(pid)
                rows.append((self.cur_un_pts[j, 0], self.cur_un_pts[j, 1], 1.0, pid * 2 + 0,
                             self.cur_pts[j, 0], self.cur_pts[j, 1], self.pts_velocity[j, 0],
                             self.pts_velocity[j, 1]))
        for j in range(len(self.ids_right)):
            pid = int(self.ids_right[j])
            if pid in pub:
                rows.append((self.cur_un_right_pts[j, 0], self.cur_un_right_pts[j, 1], 1.0,
                             pid * 2 + 1, self.cur_right_pts[j, 0], self.cur_right_pts[j, 1],
                             self.right_pts_velocity[j, 0], self.right_pts_velocity[j, 1]))
        return np.asarray(rows, np.float32).reshape(-1, 8)
