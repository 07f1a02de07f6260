"""Generates tests/golden/*.npz -- outputs of REAL OpenCV (cv2) for the third-party arithmetic
the hot path calls but the reference tree does not contain:
calcOpticalFlowPyrLK / buildOpticalFlowPyramid (feature_tracker.cpp:410,417,490,495),
findFundamentalMat (:935), circle (:30,148), convertTo(CV_8U) (event_detector.cc:260).

Run once in the build container (cv2 4.13.0):  python tests/golden/make_golden.py
The fixtures are committed; the tests never need cv2 or /root/reference at run time.
Inputs are stored next to the outputs, so the fixtures do not depend on RNG stability.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from esvio_b200 import synth  # noqa: E402
from oracle import oracle as ora  # noqa: E402  (only to render realistic time-surface inputs)


def blurred_noise(rng, h, w, shift):
    big = rng.integers(0, 256, (h + 40, w + 40)).astype(np.float32)
    big = cv2.GaussianBlur(big, (0, 0), 2.0)
    big = (big - big.min()) / (big.max() - big.min()) * 255.0
    a = big[20:20 + h, 20:20 + w]
    M = np.float32([[1, 0, shift[0]], [0, 1, shift[1]]])
    b = cv2.warpAffine(big, M, (w + 40, h + 40), flags=cv2.INTER_LINEAR)[20:20 + h, 20:20 + w]
    return np.clip(np.rint(a), 0, 255).astype(np.uint8), np.clip(np.rint(b), 0, 255).astype(np.uint8)


def lk_case(a, b, pts):
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 1, 2)
    fwd, st_f, _ = cv2.calcOpticalFlowPyrLK(a, b, pts, None, winSize=(21, 21), maxLevel=3)
    rev, st_r, _ = cv2.calcOpticalFlowPyrLK(
        b, a, fwd, pts.copy(), winSize=(21, 21), maxLevel=1,
        criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
        flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    back, st_b, _ = cv2.calcOpticalFlowPyrLK(b, a, fwd, None, winSize=(21, 21), maxLevel=3)
    return dict(a=a, b=b, pts=pts.reshape(-1, 2), fwd=fwd.reshape(-1, 2), st_f=st_f.reshape(-1),
                rev=rev.reshape(-1, 2), st_r=st_r.reshape(-1), back=back.reshape(-1, 2),
                st_b=st_b.reshape(-1))


def make_lk(rng):
    out = {}
    # (1) textured pair, 160x120: pyramid stops at 3 images (20x15 <= 21)
    a, b = blurred_noise(rng, 120, 160, (1.3, -0.7))
    pts = np.stack([rng.uniform(-2, 162, 64), rng.uniform(-2, 122, 64)], 1)
    pts[:8] = [[0.2, 0.3], [159.5, 119.4], [3.0, 60.0], [157.2, 5.1], [80.0, 0.0], [80.5, 119.9],
               [10.0, 10.0], [150.0, 110.0]]
    for k, v in lk_case(a, b, pts).items():
        out["noise_" + k] = v
    # (2) real time surfaces of the synthetic DAVIS346 stream (windows 3 -> 4, and L -> R)
    W, H = 346, 260
    s = synth.StereoEventStream(W, H, 1.0e6)
    saeL, saeR = ora.Sae(W, H), ora.Sae(W, H)
    ts = []
    for k in range(5):
        L, R, t_ref = s.stereo_window(k)
        saeL.update(*L)
        saeR.update(*R)
        ts.append((saeL.time_surface(t_ref), saeR.time_surface(t_ref)))
    prevL, curL, curR = ts[3][0], ts[4][0], ts[4][1]
    L, _, _ = s.stereo_window(4)
    sel = rng.choice(len(L[0]), 96, replace=False)
    pts = np.stack([L[0][sel], L[1][sel]], 1).astype(np.float32) + rng.uniform(-0.5, 0.5, (96, 2))
    for k, v in lk_case(prevL, curL, pts).items():
        out["ts_" + k] = v
    for k, v in lk_case(curL, curR, pts).items():
        out["stereo_" + k] = v
    # pyramid levels exactly as buildOpticalFlowPyramid produces them
    for name, img in (("noise", a), ("ts", curL), ("vga", cv2.resize(curL, (640, 480)))):
        n, levels = cv2.buildOpticalFlowPyramid(img, (21, 21), 3, withDerivatives=False)
        out[f"pyr_{name}_img"] = img
        out[f"pyr_{name}_n"] = np.int32(n + 1)
        for l, lv in enumerate(levels):
            # the Python binding hands back the level ROI without its winSize border
            out[f"pyr_{name}_l{l}"] = np.ascontiguousarray(lv)
    return out


def make_fmat(rng):
    out = {}
    cases = []

    def scene(n, outlier_frac, noise):
        # static 3-D points seen by a translating + rotating pinhole camera (f = 460)
        X = np.stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(4, 12, n)], 1)
        f, cx, cy = 460.0, 173.0, 130.0
        ang = 0.03
        R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
        t = np.array([0.25, 0.02, 0.05])
        X2 = X @ R.T + t
        p1 = np.stack([f * X[:, 0] / X[:, 2] + cx, f * X[:, 1] / X[:, 2] + cy], 1)
        p2 = np.stack([f * X2[:, 0] / X2[:, 2] + cx, f * X2[:, 1] / X2[:, 2] + cy], 1)
        p1 += rng.normal(0, noise, p1.shape)
        p2 += rng.normal(0, noise, p2.shape)
        k = int(round(outlier_frac * n))
        if k:
            idx = rng.choice(n, k, replace=False)
            p2[idx] += rng.uniform(-25, 25, (k, 2))
        return p1.astype(np.float32), p2.astype(np.float32)

    specs = [(150, 0.2, 0.2), (150, 0.05, 0.1), (150, 0.5, 0.3), (60, 0.3, 0.2), (40, 0.1, 0.3),
             (15, 0.2, 0.1), (16, 0.0, 0.05), (200, 0.35, 0.25), (300, 0.1, 0.15),
             (8, 0.0, 0.1), (9, 0.12, 0.1), (10, 0.2, 0.1), (11, 0.1, 0.2), (12, 0.25, 0.1),
             (13, 0.15, 0.1), (14, 0.2, 0.15), (14, 0.0, 0.3), (24, 0.6, 0.2)]
    for i, (n, of, nz) in enumerate(specs):
        p1, p2 = scene(n, of, nz)
        F, mask = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99)
        m = np.zeros(n, np.uint8) if mask is None else mask.reshape(-1).astype(np.uint8)
        out[f"fm{i}_p1"], out[f"fm{i}_p2"], out[f"fm{i}_mask"] = p1, p2, m
        out[f"fm{i}_ok"] = np.int32(0 if F is None else 1)
        cases.append(i)
    out["fm_cases"] = np.asarray(cases, np.int32)
    return out


def make_misc(rng):
    out = {}
    # convertTo(CV_8U) of 255*(m+1)/2 (event_detector.cc:256-260), including the empty pixel
    # The MatExpr folds to convertTo(CV_64F, alpha=127.5, beta=127.5) followed by
    # convertTo(CV_8U); cv2.normalize(NORM_MINMAX, 0..255) over data spanning exactly [-1, 1]
    # issues the same scaled convertTo, cv2.add(dtype=CV_8U) the same saturate_cast.
    m = np.concatenate([np.zeros(4), rng.uniform(-1, 1, 3000), -np.exp(-rng.uniform(0, 8, 3000)),
                        np.exp(-rng.uniform(0, 8, 3000)), [-1.0, 1.0]])
    mat = m.reshape(1, -1).astype(np.float64)
    scaled = cv2.normalize(mat, None, 0, 255, cv2.NORM_MINMAX, dtype=cv2.CV_64F)
    out["cvt_in"] = m
    out["cvt_out"] = _convert_to_u8(scaled).reshape(-1)
    # filled circles (feature_tracker.cpp:30,148): CV_64F mask, colour 255.0, thickness -1
    for r in list(range(1, 41)) + [64]:
        img = np.zeros((2 * r + 9, 2 * r + 9), np.float64)
        cv2.circle(img, (r + 4, r + 4), r, 255.0, -1)
        hw = np.full(r + 1, -1, np.int32)
        for k in range(r + 1):
            row = np.nonzero(img[r + 4 + k] == 255.0)[0]
            if len(row):
                assert row[0] + row[-1] == 2 * (r + 4) and len(row) == row[-1] - row[0] + 1
                hw[k] = (row[-1] - row[0]) // 2
            up = np.nonzero(img[r + 4 - k] == 255.0)[0]
            assert len(up) == len(row)
        out[f"disc_hw_{r}"] = hw
    # a clipped disc near the image corner
    img = np.zeros((40, 50), np.float64)
    cv2.circle(img, (3, 36), 10, 255.0, -1)
    cv2.circle(img, (48, 2), 10, 255.0, -1)
    out["disc_clip"] = (img == 255.0).astype(np.uint8)
    return out


def _convert_to_u8(m64):
    """saturate_cast<uchar>(double) == cvRound + clamp, the conversion Mat::convertTo(CV_8U)
    applies per element; cv2.add(..., dtype=CV_8U) runs the same cast on the f64 sum."""
    return cv2.add(m64, np.zeros_like(m64), dtype=cv2.CV_8U)


def main():
    rng = np.random.default_rng(20261017)
    np.savez_compressed(os.path.join(HERE, "lk_golden.npz"), **make_lk(rng))
    np.savez_compressed(os.path.join(HERE, "fmat_golden.npz"), **make_fmat(rng))
    np.savez_compressed(os.path.join(HERE, "misc_golden.npz"), **make_misc(rng))
    with open(os.path.join(HERE, "VERSIONS.txt"), "w") as f:
        f.write(f"cv2 {cv2.__version__}\nnumpy {np.__version__}\n")
    for n in ("lk_golden.npz", "fmat_golden.npz", "misc_golden.npz"):
        print(n, os.path.getsize(os.path.join(HERE, n)), "bytes")


if __name__ == "__main__":
    main()
