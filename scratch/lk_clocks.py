"""Per-phase clock64 of k_lk (needs a build with `make -C esvio_b200/csrc -B EXTRA=-DESVIO_LK_CLOCKS`)."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from esvio_b200 import frontend, _capi
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "stereo_vga_5mevs"
w, cfg, pub_div = bench.workload_cfg(wl)
cfg = dict(cfg, device_id=0, max_events_per_window=int(w["rate"]/30)+64)
if len(sys.argv) > 2:
    cfg["max_cnt"] = int(sys.argv[2])
wins = bench.gen_windows(w, 0, 7)
fe = frontend.EventFrontEnd(cfg)
L = C.CDLL(_capi.LIB_PATH)
for k in range(7):
    r = fe.track(wins[k][2], wins[k][0], wins[k][1], k % pub_div == 0)
torch.cuda.synchronize()
n = len(r["id"])
buf = np.zeros((1024, 2, 24), np.int64)
L.esvio_dbg_lk_clocks(buf.ctypes.data_as(C.c_void_p))
b = buf[:n]   # last kernel: stereo LK of window 6 (pub): fwd 4 levels + bwd 4 levels
print(f"stereo LK of the last window, {n} points; cycles (mean / max over points)")
for call, nm in ((0, "fwd"), (1, "bwd")):
    c = b[:, call]
    rows = [("phase1 loads+top region", 0, 1), ("phase2 Scharr", 1, 2), ("phase3 templates", 2, 3)]
    for name, i, j in rows:
        d = c[:, j] - c[:, i]
        print(f"  {nm} {name:26s} mean {d.mean():8.0f} max {d.max():8.0f}")
    prev = c[:, 3]
    for lv in (3, 2, 1, 0):
        ok = c[:, 6 + 4 * lv] > 0
        if not ok.any(): continue
        ens = (c[:, 5 + 4 * lv] - c[:, 4 + 4 * lv])[ok]
        newt = (c[:, 6 + 4 * lv] - c[:, 5 + 4 * lv])[ok]
        its = c[:, 7 + 4 * lv][ok]
        print(f"  {nm} level {lv}: ensure-region mean {ens.mean():7.0f} max {ens.max():7.0f} | newton mean {newt.mean():8.0f} max {newt.max():8.0f} | iterations mean {its.mean():5.1f} max {its.max():3d} | cycles/iter {newt.sum()/max(its.sum(),1):6.0f}")
    tot = c[:, 20] - c[:, 0]
    print(f"  {nm} whole call mean {tot.mean():8.0f} max {tot.max():8.0f}  ({tot.max()/1965:.1f} us at 1965 MHz)")
