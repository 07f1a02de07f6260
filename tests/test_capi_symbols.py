"""CPU suite: the C-ABI library builds, loads and exports every symbol include/esvio_fe.h
declares; entry points that need no GPU behave (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from esvio_b200 import _capi


def _declared_functions():
    src = open(_capi.HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(esvio_fe_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_and_binding_agree():
    declared = _declared_functions()
    assert len(declared) >= 25
    assert sorted(_capi.SYMBOLS) == declared


def test_library_exports_every_declared_symbol(capi):
    lib = C.CDLL(capi.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} missing from libesvio_fe.so"
    assert capi.lib().esvio_fe_abi_version() == 3


def test_strerror_and_default_config(capi):
    L = capi.lib()
    msgs = [L.esvio_fe_strerror(i).decode() for i in range(6)]
    assert msgs[0] == "ok" and len(set(msgs)) == 6
    c = capi.Config()
    L.esvio_fe_default_config(C.byref(c), 346, 260)
    # values common to every shipped config
    assert (c.width, c.height, c.max_cnt, c.min_dist, c.flow_back) == (346, 260, 150, 10, 1)
    assert (c.f_threshold, c.ts_lk_threshold, c.decay_ms) == (1.0, 128.0, 20.0)
    assert c.feature_filter_threshold == 0.01 and c.focal_length == 460.0 and c.use_ransac == 1
    assert C.sizeof(capi.Config) == 256 or C.sizeof(capi.Config) % 8 == 0


def test_create_rejects_bad_config_or_missing_gpu(capi):
    L = capi.lib()
    c = capi.Config()
    L.esvio_fe_default_config(C.byref(c), 346, 260)
    h = C.c_void_p()
    c.max_cnt = 0
    assert L.esvio_fe_create(C.byref(c), C.byref(h)) == capi.EINVAL and not h.value
    c.max_cnt = 150
    c.median_blur_kernel_size = 8   # medianBlur(17): beyond what the kernel stages -> refused
    assert L.esvio_fe_create(C.byref(c), C.byref(h)) == capi.EINVAL
    c.median_blur_kernel_size = 0
    assert L.esvio_fe_create(None, C.byref(h)) == capi.EINVAL
    import torch
    if not torch.cuda.is_available():
        # there is no CPU fallback: creation fails loudly without a device
        assert L.esvio_fe_create(C.byref(c), C.byref(h)) == capi.ENODEV and not h.value


def test_null_handle_is_einval(capi):
    L = capi.lib()
    assert L.esvio_fe_reset(None) == capi.EINVAL
    assert L.esvio_fe_track_wait(None, None) == capi.EINVAL
    n = C.c_int64()
    assert L.esvio_fe_kernel_launches(None, C.byref(n)) == capi.EINVAL


def test_header_is_plain_c99(tmp_path):
    """The boundary is a C ABI: include/esvio_fe.h must compile as C99 with no C++ or torch types
    (a cgo / JNI / ctypes binding sees exactly this)."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "esvio_fe.h"\n'
                   "int main(void) { esvio_fe_config c; esvio_tracks t; esvio_events e; esvio_motion m;\n"
                   "  (void)c; (void)t; (void)e; (void)m; return ESVIO_FE_ABI_VERSION == 3 ? 0 : 1; }\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I",
                           os.path.join(root, "include"), "-c", str(src), "-o", str(tmp_path / "hdr.o")])
