"""std_sort_order (the libstdc++ std::sort replay inside k_select / k_image_set_mask) and the
selection kernel with enough tracks to leave the insertion-sort-only case, small enough for
compute-sanitizer:  compute-sanitizer --tool {memcheck,racecheck,synccheck} python scratch/sanitize_sort.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esvio_b200 import frontend, synth

W, H = 346, 260
cfg = synth.default_config(W, H, max_cnt=300, max_events_per_window=1 << 15)
fe = frontend.EventFrontEnd(cfg)
rng = np.random.default_rng(1)
for n, span, depth in ((40, 3, -1), (150, 5, -1), (150, 5, 1), (700, 2, -1), (1024, 1000, 0), (1024, 1, -1)):
    key = rng.integers(0, span, n).astype(np.int32)
    o = fe.stage_sort_order(key, depth)
    assert sorted(o.tolist()) == list(range(n)) and np.all(np.diff(key[o]) <= 0)
s = synth.StereoEventStream(W, H, 6.0e5)
L, R, t = s.stereo_window(0)
fe.stage_update(t, L, R)
for n in (150, 290):
    pts = np.stack([rng.uniform(1, W - 2, n), rng.uniform(1, H - 2, n)], 1).astype(np.float32)
    po, io, co, kept = fe.stage_select(L, pts, np.arange(n, dtype=np.int32), rng.integers(1, 4, n).astype(np.int32))
    print("select", n, kept, len(po))
fe.close()
print("sanitize_sort done")
