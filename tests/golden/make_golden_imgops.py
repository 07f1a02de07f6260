"""Generates tests/golden/imgops_golden.npz -- outputs of REAL OpenCV (cv2) for the optional
image conditioning of the time surface:
  cv::medianBlur(ts, ts, 2k+1)                          event_detector.cc:262-264
  cv::createCLAHE()->apply + cv::normalize(0,255,MINMAX) feature_tracker.cpp:375-382
Run once in the build container (cv2 4.13.0):  python tests/golden/make_golden_imgops.py
Inputs are stored next to the outputs.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from esvio_b200 import synth  # noqa: E402
from oracle import oracle as ora  # noqa: E402  (only to render realistic time-surface inputs)


def time_surfaces(W, H, rate, windows=3):
    s = synth.StereoEventStream(W, H, rate)
    sae = ora.Sae(W, H)
    for k in range(windows):
        L, _, t_ref = s.stereo_window(k)
        sae.update(*L)
    return sae.time_surface(t_ref)


def main():
    rng = np.random.default_rng(20260117)
    imgs = {
        "ts346": time_surfaces(346, 260, 1.0e6),
        "ts640": time_surfaces(640, 480, 5.0e6),
        "noise173": rng.integers(0, 256, (130, 173)).astype(np.uint8),
        "noise160": rng.integers(0, 256, (120, 160)).astype(np.uint8),
        "flat_w8": np.full((96, 104), 128, np.uint8),        # width % 8 == 0, constant image
        "ramp_h8": (np.arange(100 * 96).reshape(96, 100) % 251).astype(np.uint8),  # height % 8 == 0 only
        "lowrange": rng.integers(120, 136, (130, 173)).astype(np.uint8),
    }
    out = {}
    clahe = cv2.createCLAHE()
    assert clahe.getClipLimit() == 40.0 and clahe.getTilesGridSize() == (8, 8)
    for name, im in imgs.items():
        out[name] = im
        for k in (1, 2, 3):
            out[f"{name}_median{2 * k + 1}"] = cv2.medianBlur(im, 2 * k + 1)
        c = clahe.apply(im)
        out[name + "_clahe"] = c
        out[name + "_clahe_norm"] = cv2.normalize(c, None, 0, 255, cv2.NORM_MINMAX)
        out[name + "_norm"] = cv2.normalize(im, None, 0, 255, cv2.NORM_MINMAX)
    np.savez_compressed(os.path.join(HERE, "imgops_golden.npz"), **out)
    print("wrote imgops_golden.npz:", len(out), "arrays, cv2", cv2.__version__)


if __name__ == "__main__":
    main()
