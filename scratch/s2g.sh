#!/bin/bash
T=${1:-s2g}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
for w in stereo_vga_5mevs stereo_vga_10mevs stereo_davis346_1mevs; do python scratch/stage_times.py $w 40 2>&1 | tail -1; done
python bench.py --steps 30 --warmup 6 --no-cpu --no-frames --no-secondary --no-rigid > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f sync_ms %.3f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["sync_call_ms_per_step"], d["ms_per_step"]))
print("stage_ms", {k: round(v*1e3,1) for k,v in d["stage_ms"].items()})
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "batched", d.get("batched",{}).get("value"))
PY
