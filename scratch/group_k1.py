"""Batched k_sae_update_ts in isolation: group of S streams, synchronous windows."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esvio_b200 import frontend, synth
name = sys.argv[1] if len(sys.argv) > 1 else "stereo_davis346_1mevs"
w = synth.WORKLOADS[name]
cfg = synth.default_config(w["width"], w["height"], max_cnt=w["max_cnt"], min_dist=w["min_dist"], use_ransac=1)
cfg["max_events_per_window"] = int(w["rate"] / 30) + 1024
for S in [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]:
    g = frontend.EventFrontEndGroup(cfg, S)
    m0 = g.member(0)
    streams = [synth.StereoEventStream(w["width"], w["height"], w["rate"], stream=i) for i in range(S)]
    ms = []
    nev = 0
    for k in range(16):
        ws = [st.stereo_window(k) for st in streams]
        g.submit([x[2] for x in ws], [x[0] for x in ws], [x[1] for x in ws], [k % 2 == 0] * S)
        g.wait(unpack=False)
        if k >= 6:
            ms.append(g.sae_ts_ms()); nev = sum(len(x[0][0]) + len(x[1][0]) for x in ws)
    alg = S * 2 * 17 * w["width"] * w["height"] + 45 * nev
    k1 = np.mean(ms) * 1e-3
    print(f"{name} S={S} k1_us={k1*1e6:.1f} min={min(ms)*1e3:.1f} alg_MB={alg/1e6:.1f} GB/s={alg/k1/1e9:.0f} frac={alg/k1/1e9/6538:.3f}")
    g.close()
