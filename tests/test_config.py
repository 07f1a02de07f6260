"""CPU suite: the configuration-set loader (VERDICT r1 "shipped-config coverage").
esvio_b200/config.py and its C++ twin include/esvio_fe_config.hpp read an OpenCV-YAML
`es*io.yaml` plus the two camodocal PINHOLE calibrations it names, the way
readParameters_event (feature_tracker/src/parameters.cpp:183-282) does."""
import glob
import os
import subprocess
import textwrap

import numpy as np
import pytest

from esvio_b200 import config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "config_dump")
REF_CONFIG = "/root/reference/config"

MAIN = """\
%YAML:1.0

#common parameters
imu_topic: "/imu/data"   # a comment with a : colon
event_left_topic: "/davis/left/events"
event_left_calib: "ev0.yaml"
event_right_calib: "ev1.yaml"
event_width: 346
event_height: 260
extrinsicRotation: !!opencv-matrix
   rows: 3
   cols: 3
   dt: d
   data: [-0.99973298, -0.00994674, 0.02085725,
           0.01003579, -0.99994095, 0.00416910,
           0.02081454, 0.00437730, 0.99977377]
Trl_event: !!opencv-matrix
   rows: 3
   cols: 1
   dt: d
   data: [-0.04372224 , 0.00101557, -0.01337267]
ignore_polarity: 0 #true 1 false 0;
decay_ms: 20 #20
median_blur_kernel_size: 0
feature_filter_threshold: 0.01
TS_LK_threshold: 128.0
max_cnt: 200
min_dist: 20 #10
freq: 0
F_threshold: 1.0
flow_back: 1
equalize: 1
Do_motion_correction: 1
"""
CAM = """\
%YAML:1.0
---
model_type: PINHOLE
camera_name: camera
image_width: 346
image_height: 260
distortion_parameters:
   k1: {k1}
   k2: 0.15393049046270296
   p1: 7.642434980998821e-05
   p2: -0.0019042695753031854
projection_parameters:
   fx: 249.69341447817564
   fy: 248.41625664694038
   cx: 176.74240257052816
   cy: 129.47631010746218
"""


@pytest.fixture()
def cfg_dir(tmp_path):
    (tmp_path / "esvio.yaml").write_text(MAIN)
    (tmp_path / "ev0.yaml").write_text(CAM.format(k1=-0.3794794654640921))
    (tmp_path / "ev1.yaml").write_text(CAM.format(k1=-0.25))
    return str(tmp_path)


def _cpp(config_file):
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "config_dump.cpp"), "-o", EXE])
    r = subprocess.run([EXE, config_file], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    return dict(l.split(" ", 1) for l in r.stdout.splitlines())


def _same(cpp, cfg, node):
    for k in ("width", "height", "max_cnt", "min_dist", "flow_back", "equalize", "ignore_polarity",
              "median_blur_kernel_size", "do_motion_correction"):
        assert int(cpp[k]) == cfg[k], k
    for k in ("f_threshold", "ts_lk_threshold", "decay_ms", "feature_filter_threshold", "focal_length"):
        assert float(cpp[k]) == cfg[k], k
    assert int(cpp["freq"]) == node["freq"]
    for i in range(2):
        got = [float(v) for v in cpp[f"cam{i}"].split()]
        ref = [cfg["cam"][i][k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")]
        assert got == ref, (i, got, ref)
    assert cpp["event_left_topic"] == node["event_left_topic"]


def test_reader_on_the_opencv_yaml_dialect(cfg_dir):
    y = config.read_opencv_yaml(os.path.join(cfg_dir, "esvio.yaml"))
    assert y["imu_topic"] == "/imu/data" and y["max_cnt"] == 200 and y["decay_ms"] == 20
    assert y["extrinsicRotation"].shape == (3, 3) and y["extrinsicRotation"][2, 2] == 0.99977377
    assert np.array_equal(y["Trl_event"].ravel(), [-0.04372224, 0.00101557, -0.01337267])
    cfg, node = config.read_parameters_event(config.find_config(cfg_dir))
    assert (cfg["width"], cfg["height"], cfg["max_cnt"], cfg["min_dist"]) == (346, 260, 200, 20)
    assert cfg["equalize"] == 1 and cfg["do_motion_correction"] == 1 and cfg["focal_length"] == 460.0
    assert node["freq"] == 100                       # freq 0 -> 100 (parameters.cpp:277-278)
    assert cfg["cam"][0]["k1"] == -0.3794794654640921 and cfg["cam"][1]["k1"] == -0.25
    assert cfg["cam"][0]["p1"] == 7.642434980998821e-05
    _same(_cpp(os.path.join(cfg_dir, "esvio.yaml")), cfg, node)
    # the dict is what the ABI's config struct takes
    from esvio_b200 import _capi
    if os.path.exists(_capi.LIB_PATH):
        from esvio_b200 import frontend
        c = frontend.make_config(cfg)
        assert (c.width, c.max_cnt, c.min_dist, c.equalize) == (346, 200, 20, 1) and c.cam[1].k1 == -0.25


def test_non_pinhole_calibration_is_refused(cfg_dir):
    with open(os.path.join(cfg_dir, "ev1.yaml"), "w") as f:
        f.write(CAM.format(k1=0.1).replace("PINHOLE", "MEI"))
    with pytest.raises(ValueError):
        config.read_parameters_event(os.path.join(cfg_dir, "esvio.yaml"))


@pytest.mark.skipif(not os.path.isdir(REF_CONFIG), reason="/root/reference/config absent")
def test_every_shipped_set_loads_with_the_surveyed_values():
    """SURVEY.md 5.6, line by line, and the C++ twin on the same files."""
    expect = {  # set: (W, H, max_cnt, min_dist, freq, equalize, Do_motion_correction)
        "esvio": (346, 260, 150, 10, 15, 0, 0), "esio": (346, 260, 150, 10, 15, 0, 0),
        "esvio_DSEC": (640, 480, 100, 30, 10, 0, 0), "esio_DSEC": (640, 480, 300, 10, 15, 1, 0),
        "esvio_VECtor": (640, 480, 150, 10, 10, 0, 1),
        "esvio_VECtor_small_scale": (640, 480, 150, 10, 10, 0, 1),
        "esvio_ecmd": (640, 480, 200, 20, 10, 0, 0), "esvio_mvsec_flying": (346, 260, 150, 10, 15, 0, 1)}
    dirs = sorted(d for d in glob.glob(os.path.join(REF_CONFIG, "es*")) if os.path.isdir(d))
    assert {os.path.basename(d) for d in dirs} == set(expect)
    for d in dirs:
        f = config.find_config(d)
        cfg, node = config.read_parameters_event(f)
        got = (cfg["width"], cfg["height"], cfg["max_cnt"], cfg["min_dist"], node["freq"], cfg["equalize"],
               cfg["do_motion_correction"])
        assert got == expect[os.path.basename(d)], (d, got)
        assert (cfg["decay_ms"], cfg["feature_filter_threshold"], cfg["ts_lk_threshold"], cfg["f_threshold"],
                cfg["flow_back"], cfg["ignore_polarity"], cfg["median_blur_kernel_size"]) == (20, 0.01, 128.0, 1.0, 1, 0, 0)
        assert all(c["fx"] > 100 and c["fy"] > 100 for c in cfg["cam"])
        _same(_cpp(f), cfg, node)
