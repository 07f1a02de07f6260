"""Loader for the reference's configuration sets (`config/<set>/es*io.yaml` plus the two
`event{0,1}_esvio.yaml` calibrations), i.e. what `readParameters_event`
(/root/reference/feature_tracker/src/parameters.cpp:183-282) and
`FeatureTracker::stereo_readIntrinsicParameter` (feature_tracker.cpp:963-976) read at node
start-up, turned into the plain dict `frontend.make_config` / `esvio_fe_config` takes.

The files are OpenCV `FileStorage` YAML 1.0 (`%YAML:1.0` header, `!!opencv-matrix` maps, flow
sequences that may span lines); the reader below covers that dialect with no dependency
(the C++ twin is include/esvio_fe_config.hpp).  Like `cv::FileNode`, a missing numeric key reads
as 0 (e.g. `fx` in every shipped set, parameters.cpp:221).
"""
from __future__ import annotations

import os

import numpy as np


def _strip_comment(line: str) -> str:
    out, quote = [], None
    for ch in line:
        if quote:
            if ch == quote:
                quote = None
        elif ch in "\"'":
            quote = ch
        elif ch == "#":
            break
        out.append(ch)
    return "".join(out).rstrip()


def _scalar(tok: str):
    tok = tok.strip()
    if len(tok) >= 2 and tok[0] == tok[-1] and tok[0] in "\"'":
        return tok[1:-1]
    try:
        return int(tok)
    except ValueError:
        pass
    try:
        return float(tok)
    except ValueError:
        return tok


def read_opencv_yaml(path: str) -> dict:
    """Parse an OpenCV-YAML file into nested dicts; `!!opencv-matrix` maps become numpy arrays."""
    root: dict = {}
    stack = [(-1, root)]          # (indent of the map's keys' parent, map)
    pending = None                # (map, key, text) of a flow sequence still open
    with open(path) as f:
        for raw in f:
            line = _strip_comment(raw.rstrip("\n"))
            if pending is not None:
                m, key, text = pending
                text += " " + line.strip()
                if "]" in line:
                    m[key] = [_scalar(t) for t in text[text.index("[") + 1: text.rindex("]")].split(",") if t.strip()]
                    pending = None
                else:
                    pending = (m, key, text)
                continue
            if not line.strip() or line.lstrip().startswith("%") or line.strip() == "---":
                continue
            indent = len(line) - len(line.lstrip())
            body = line.strip()
            if ":" not in body:
                continue
            key, _, val = body.partition(":")
            key, val = key.strip(), val.strip()
            while len(stack) > 1 and indent <= stack[-1][0]:
                stack.pop()
            cur = stack[-1][1]
            if val == "" or val.startswith("!!"):
                child: dict = {}
                if val.startswith("!!"):
                    child["__tag__"] = val
                cur[key] = child
                stack.append((indent, child))
            elif val.startswith("["):
                if "]" in val:
                    cur[key] = [_scalar(t) for t in val[1: val.rindex("]")].split(",") if t.strip()]
                else:
                    pending = (cur, key, val)
            else:
                cur[key] = _scalar(val)

    def finish(m):
        for k, v in list(m.items()):
            if isinstance(v, dict):
                if v.get("__tag__") == "!!opencv-matrix":
                    dt = np.float64 if v.get("dt", "d") == "d" else np.float32
                    m[k] = np.asarray(v["data"], dt).reshape(int(v["rows"]), int(v["cols"]))
                else:
                    finish(v)
        return m

    return finish(root)


def read_pinhole_yaml(path: str) -> dict:
    """camodocal PINHOLE calibration (PinholeCamera::Parameters::readFromYamlFile,
    camera_model/src/camera_models/PinholeCamera.cc:150-191): fx fy cx cy k1 k2 p1 p2."""
    y = read_opencv_yaml(path)
    model = str(y.get("model_type", "PINHOLE"))
    if model != "PINHOLE":
        raise ValueError(f"{path}: model_type {model}; the event front-end lifts with the PINHOLE "
                         "model only (every shipped event calibration is PINHOLE)")
    d, p = y.get("distortion_parameters", {}), y.get("projection_parameters", {})
    out = {k: float(d.get(k, 0.0)) for k in ("k1", "k2", "p1", "p2")}
    out.update({k: float(p.get(k, 0.0)) for k in ("fx", "fy", "cx", "cy")})
    out["image_width"], out["image_height"] = int(y.get("image_width", 0)), int(y.get("image_height", 0))
    return out


def read_parameters_event(config_file: str, esvio_folder: str | None = None):
    """readParameters_event: (front-end config dict, node dict).  `esvio_folder` is the ROS
    param of the same name that prefixes the calibration files (parameters.cpp:192,243-244);
    default: the directory of `config_file`, which is what the shipped launch files pass."""
    y = read_opencv_yaml(config_file)
    folder = esvio_folder or os.path.dirname(os.path.abspath(config_file))
    num = lambda k: y.get(k, 0) if not isinstance(y.get(k, 0), str) else 0   # noqa: E731
    cams = [read_pinhole_yaml(os.path.join(folder, str(y[k])))
            for k in ("event_left_calib", "event_right_calib")]
    cfg = dict(
        width=int(num("event_width")), height=int(num("event_height")),
        max_cnt=int(num("max_cnt")), min_dist=int(num("min_dist")), flow_back=int(num("flow_back")),
        equalize=int(num("equalize")), f_threshold=float(num("F_threshold")),
        ts_lk_threshold=float(num("TS_LK_threshold")), decay_ms=float(num("decay_ms")),
        ignore_polarity=int(num("ignore_polarity")),
        median_blur_kernel_size=int(num("median_blur_kernel_size")),
        feature_filter_threshold=float(num("feature_filter_threshold")),
        do_motion_correction=int(num("Do_motion_correction")),
        focal_length=460.0,                         # FOCAL_LENGTH = 460 (parameters.cpp:274)
        use_ransac=1,
        cam=tuple({k: c[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")} for c in cams),
    )
    freq = int(num("freq"))
    node = dict(freq=freq if freq != 0 else 100,    # parameters.cpp:277-278
                show_track=int(num("show_track")), fisheye=int(num("fisheye")),
                max_cnt_img=int(num("max_cnt_img")), min_dist_img=int(num("min_dist_img")),
                image_width=int(num("image_width")), image_height=int(num("image_height")),
                event_left_topic=str(y.get("event_left_topic", "")),
                event_right_topic=str(y.get("event_right_topic", "")),
                imu_topic=str(y.get("imu_topic", "")))
    return cfg, node


def find_config(config_dir: str) -> str:
    """The `es*io.yaml` of a shipped set directory (config/esvio, config/esio_DSEC, ...)."""
    for name in ("esvio.yaml", "esio.yaml"):
        p = os.path.join(config_dir, name)
        if os.path.exists(p):
            return p
    raise FileNotFoundError(f"no esvio.yaml / esio.yaml under {config_dir}")
