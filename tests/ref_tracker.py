"""pytest glue for oracle/ref_tracker.py (the reference's own FeatureTracker behind ctypes)."""
import pytest

from oracle.ref_tracker import KEYS_L, KEYS_R, RefTracker, _p, load  # noqa: F401


def load_ref_lib():
    L = load()
    if L is None:
        pytest.skip("oracle/_ref/libesvio_ref_ft.so not built and /root/reference absent")
    return L
