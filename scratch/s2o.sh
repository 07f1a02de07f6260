#!/bin/bash
T=${1:-s2o}
for nw in ${NWS:-4 2 1}; do
echo "== NW $nw"
ESVIO_LK_NW=$nw python -m pytest tests -m gpu -x -q -k "lk_matches or teacher_forced" 2>&1 | tail -1
ESVIO_LK_NW=$nw ESVIO_FE_LIB=$PWD/scratch/variants/libesvio_fe_clk.so python scratch/lk_clocks.py stereo_vga_5mevs 2>&1 | grep -E "cycles/iter|whole call"
ESVIO_LK_NW=$nw python bench.py --steps 40 --warmup 6 --no-cpu --no-frames --no-secondary --no-rigid > gpurun_out/${T}_bench$nw.json 2> gpurun_out/${T}_bench$nw.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench$nw.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f sync_ms %.3f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["sync_call_ms_per_step"], d["ms_per_step"]))
print("stage_ms", {k: round(v*1e3,1) for k,v in d["stage_ms"].items()})
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "batched", d.get("batched",{}).get("value"))
PY
done
