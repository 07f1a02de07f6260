"""ctypes binding of libesvio_fe.so (include/esvio_fe.h).

The library is built in-tree (esvio_b200/csrc/Makefile, sm_100a only).  There is no CPU
implementation behind this module: loading fails loudly when the shared object is missing
and esvio_fe_create fails with ENODEV when no Blackwell GPU is visible.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("ESVIO_FE_LIB") or os.path.join(CSRC, "libesvio_fe.so")  # override: instrumented scratch builds
HEADER = os.path.join(os.path.dirname(_HERE), "include", "esvio_fe.h")

NUM_STAGES = 9
STAGE_NAMES = ("h2d", "bin_events", "sae_update_ts", "pyramid", "corner_flags", "lk_temporal", "select", "lk_stereo",
               "d2h")

OK, EINVAL, ENODEV, ECUDA, ECAPACITY, ESTATE = range(6)


class Pinhole(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")]


class Config(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("max_cnt", C.c_int32),
        ("min_dist", C.c_int32), ("flow_back", C.c_int32), ("equalize", C.c_int32),
        ("f_threshold", C.c_double), ("ts_lk_threshold", C.c_double), ("decay_ms", C.c_double),
        ("ignore_polarity", C.c_int32), ("median_blur_kernel_size", C.c_int32),
        ("feature_filter_threshold", C.c_double), ("do_motion_correction", C.c_int32),
        ("focal_length", C.c_double), ("cam", Pinhole * 2), ("device_id", C.c_int32),
        ("max_events_per_window", C.c_int32), ("use_ransac", C.c_int32),
        ("reserved0", C.c_int32),
        ("mc_fx", C.c_double), ("mc_fy", C.c_double), ("mc_cx", C.c_double), ("mc_cy", C.c_double),
        ("reserved", C.c_int32 * 6),
    ]


class Motion(C.Structure):
    """esvio_motion: the fields of Motion_correction_value the SAE update reads."""
    _fields_ = [("state_v", C.c_double * 3), ("v_pre", C.c_float * 3), ("accel", C.c_float * 3),
                ("omega", C.c_float * 3), ("t1", C.c_double)]


class Events(C.Structure):
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("t", C.c_void_p), ("p", C.c_void_p),
                ("aos", C.c_void_p), ("n", C.c_size_t), ("on_device", C.c_int32),
                ("flags", C.c_int32)]


EVENTS_STEREO_BLOCK = 1


class Stats(C.Structure):
    _fields_ = [("n_events", C.c_int32 * 2), ("n_dropped", C.c_int32 * 2),
                ("n_prev", C.c_int32), ("n_after_temporal", C.c_int32),
                ("n_after_ransac", C.c_int32), ("n_after_mask", C.c_int32), ("n_new", C.c_int32),
                ("n_corner_flags", C.c_int32), ("ransac_iters", C.c_int32),
                ("reserved", C.c_int32 * 5)]


_pf = C.POINTER(C.c_float)
_pi = C.POINTER(C.c_int32)


class Tracks(C.Structure):
    _fields_ = [("capacity", C.c_int32), ("n_left", C.c_int32), ("id", _pi), ("track_cnt", _pi),
                ("u", _pf), ("v", _pf), ("un_x", _pf), ("un_y", _pf), ("vx", _pf), ("vy", _pf),
                ("n_right", C.c_int32), ("id_right", _pi), ("ru", _pf), ("rv", _pf),
                ("run_x", _pf), ("run_y", _pf), ("rvx", _pf), ("rvy", _pf), ("stats", Stats)]


# every symbol include/esvio_fe.h declares: name -> (restype, argtypes)
_H = C.c_void_p
SYMBOLS = {
    "esvio_fe_abi_version": (C.c_int, []),
    "esvio_fe_default_config": (None, [C.POINTER(Config), C.c_int32, C.c_int32]),
    "esvio_fe_create": (C.c_int, [C.POINTER(Config), C.POINTER(_H)]),
    "esvio_fe_destroy": (None, [_H]),
    "esvio_fe_reset": (C.c_int, [_H]),
    "esvio_fe_strerror": (C.c_char_p, [C.c_int]),
    "esvio_fe_last_error": (C.c_char_p, [_H]),
    "esvio_fe_track": (C.c_int, [_H, C.c_double, C.POINTER(Events), C.POINTER(Events), C.c_int32,
                                 C.POINTER(Tracks)]),
    "esvio_fe_track_submit": (C.c_int, [_H, C.c_double, C.POINTER(Events), C.POINTER(Events),
                                        C.c_int32]),
    "esvio_fe_track_wait": (C.c_int, [_H, C.POINTER(Tracks)]),
    "esvio_fe_track_mc": (C.c_int, [_H, C.c_double, C.POINTER(Events), C.POINTER(Events), C.c_int32,
                                    C.POINTER(Motion), C.POINTER(Tracks)]),
    "esvio_fe_track_submit_mc": (C.c_int, [_H, C.c_double, C.POINTER(Events), C.POINTER(Events),
                                           C.c_int32, C.POINTER(Motion)]),
    "esvio_fe_group_create": (C.c_int, [C.POINTER(Config), C.c_int32, C.POINTER(_H)]),
    "esvio_fe_group_destroy": (None, [_H]),
    "esvio_fe_group_reset": (C.c_int, [_H]),
    "esvio_fe_group_member": (_H, [_H, C.c_int32]),
    "esvio_fe_group_track": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(Events), C.POINTER(Events),
                                       C.POINTER(C.c_int32), C.POINTER(Tracks)]),
    "esvio_fe_group_track_submit": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(Events),
                                              C.POINTER(Events), C.POINTER(C.c_int32)]),
    "esvio_fe_group_track_wait": (C.c_int, [_H, C.POINTER(Tracks)]),
    "esvio_fe_group_kernel_launches": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "esvio_fe_group_sae_ts_ms": (C.c_int, [_H, _pf]),
    "esvio_fe_time_surface": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_size_t]),
    "esvio_fe_soa_layout": (None, [C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "esvio_fe_soa_layout_stereo": (None, [C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                          C.POINTER(C.c_size_t)]),
    "esvio_fe_host_alloc": (C.c_void_p, [C.c_size_t]),
    "esvio_fe_host_free": (None, [C.c_void_p]),
    "esvio_fe_device_alloc": (C.c_int, [_H, C.c_size_t, C.POINTER(C.c_void_p)]),
    "esvio_fe_device_free": (C.c_int, [_H, C.c_void_p]),
    "esvio_fe_copy_to_device": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_size_t]),
    "esvio_fe_result_device_ptr": (C.c_int, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "esvio_fe_result_acquire": (C.c_int, [_H, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "esvio_fe_result_release": (C.c_int, [_H, C.c_void_p]),
    "esvio_fe_stream": (C.c_int, [_H, C.POINTER(C.c_void_p)]),
    "esvio_fe_split_image_submit": (C.c_int, [_H, C.c_double, C.POINTER(Events), C.c_void_p,
                                              C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "esvio_fe_split_right_buffer": (C.c_int, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "esvio_fe_track_submit_split": (C.c_int, [_H, C.c_double, C.POINTER(Events), C.c_int32,
                                              C.c_void_p]),
    "esvio_fe_track_image": (C.c_int, [_H, C.c_double, C.c_void_p, C.c_size_t, C.c_void_p,
                                       C.c_size_t, C.c_int32, C.POINTER(Tracks)]),
    "esvio_fe_track_image_submit": (C.c_int, [_H, C.c_double, C.c_void_p, C.c_size_t, C.c_void_p,
                                              C.c_size_t, C.c_int32]),
    "esvio_fe_stage_good_features": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int32, C.c_double,
                                               C.c_void_p, C.c_int32, _pi, C.c_void_p]),
    "esvio_fe_set_profiling": (C.c_int, [_H, C.c_int32]),
    "esvio_fe_get_stage_ms": (C.c_int, [_H, _pf]),
    "esvio_fe_get_stage_marks": (C.c_int, [_H, _pf]),
    "esvio_fe_pipeline_depth": (C.c_int, []),
    "esvio_fe_state_device_ptrs": (C.c_int, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "esvio_fe_shard_merge_max": (C.c_int, [_H, C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_size_t, C.c_void_p]),
    "esvio_fe_shard_event_stage": (C.c_int, [_H, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esvio_fe_shard_corner_candidates": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "esvio_fe_shard_sizes": (C.c_int, [_H, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "esvio_fe_shard_images": (C.c_int, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "esvio_fe_external_buffers": (C.c_int, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "esvio_fe_track_submit_external": (C.c_int, [_H, C.c_double, C.c_int32, C.c_int32, C.c_void_p]),
    "esvio_fe_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "esvio_fe_comm_init": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32]),
    "esvio_fe_comm_attach": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32]),
    "esvio_fe_allgather_tracks": (C.c_int, [_H]),
    "esvio_fe_gathered_tracks": (C.c_int, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]),
    "esvio_fe_stage_set_tracks": (C.c_int, [_H, C.c_double, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "esvio_fe_kernel_launches": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "esvio_fe_get_sae": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_void_p]),
    "esvio_fe_stage_update": (C.c_int, [_H, C.c_double, C.POINTER(Events), C.POINTER(Events)]),
    "esvio_fe_stage_update_mc": (C.c_int, [_H, C.c_double, C.POINTER(Events), C.POINTER(Events),
                                           C.POINTER(Motion)]),
    "esvio_fe_stage_motion_correct": (C.c_int, [_H, C.POINTER(Motion), C.c_void_p, C.c_int32,
                                                C.c_void_p]),
    "esvio_fe_stage_corner_flags": (C.c_int, [_H, C.POINTER(Events), C.c_int32, C.c_void_p]),
    "esvio_fe_get_pyramid_level": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_void_p, _pi, _pi]),
    "esvio_fe_stage_lk": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.c_void_p, C.c_int32, C.c_int32]),
    "esvio_fe_stage_condition": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "esvio_fe_stage_fmat_mask": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int32, C.c_double,
                                           C.c_void_p, _pi]),
    "esvio_fe_stage_select": (C.c_int, [_H, C.POINTER(Events), C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, _pi, C.c_void_p, C.c_void_p, C.c_void_p, _pi]),
    "esvio_fe_stage_sort_order": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "esvio_fe_stage_undistort": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
}


def build(force: bool = False) -> str:
    """Compile libesvio_fe.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    args = ["make", "-C", CSRC, "-s", "-j8"] + (["-B"] if force else [])
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C esvio_b200/csrc` "
                "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class FrontEndError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        msg = lib().esvio_fe_strerror(status).decode()
        super().__init__(f"{where}: {msg}" + (f" ({detail})" if detail else ""))
