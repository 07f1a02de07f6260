"""Phase clocks of k_select (build with EXTRA=-DESVIO_LK_CLOCKS)."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from esvio_b200 import frontend, _capi
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "stereo_vga_5mevs"
w, cfg, pub_div = bench.workload_cfg(wl)
cfg = dict(cfg, device_id=0, max_events_per_window=int(w["rate"]/30)+64)
wins = bench.gen_windows(w, 0, 10)
fe = frontend.EventFrontEnd(cfg)
L = C.CDLL(_capi.LIB_PATH)
names = ["zero mask + rank", "conflict matrix", "greedy kept", "compact + fill discs", "feature walk", "append + snapshot"]
for k in range(10):
    r = fe.track(wins[k][2], wins[k][0], wins[k][1], k % pub_div == 0)
    if k % pub_div == 0:
        torch.cuda.synchronize()
        b = np.zeros(16, np.int64)
        L.esvio_dbg_select_clocks(b.ctypes.data_as(C.c_void_p))
        d = np.diff(b[:7])
        print(f"window {k}: stats {r['stats']} total {b[6]-b[0]} cycles = {(b[6]-b[0])/1965:.1f} us | " + ", ".join(f"{n} {int(v)}" for n, v in zip(names, d)) + f" | chunks {b[10]} load cycles {b[9]} rounds-with-free {b[8]} accepted {b[11]}")
