"""The header-only C++ mirror of FeatureTracker (include/esvio_fe_adapter.hpp) compiles with
plain g++ against the C ABI and behaves: on a CPU box creation fails loudly (no CPU fallback),
on a B200 it tracks and packs PointCloud rows that satisfy the estimator's decode rules."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "adapter_smoke")


def _build(capi):
    csrc = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "adapter_smoke.cpp"), "-o", EXE,
                           "-L", csrc, "-lesvio_fe", f"-Wl,-rpath,{csrc}"])


def test_adapter_compiles_and_fails_loudly_without_gpu(capi):
    import torch
    _build(capi)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_adapter_tracks_on_gpu(capi):
    _build(capi)
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "cloud rows" in r.stdout
