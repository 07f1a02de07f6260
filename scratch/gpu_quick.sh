#!/bin/bash
# quick GPU check: parity tests + timeline + short bench
python -m pytest tests -m gpu -x -q > gpurun_out/$1_pytest.log 2>&1; tail -3 gpurun_out/$1_pytest.log
python scratch/timeline.py stereo_vga_5mevs > gpurun_out/$1_timeline.txt 2>&1; tail -9 gpurun_out/$1_timeline.txt
python bench.py --steps 20 --warmup 5 --no-cpu --no-frames --no-secondary > gpurun_out/$1_bench.json 2> gpurun_out/$1_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/$1_bench.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f sync_ms %.3f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["sync_call_ms_per_step"], d["ms_per_step"]))
print("stage_ms", {k: round(v*1e3,1) for k,v in d["stage_ms"].items()})
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "rigid e2e", d.get("rigid_scene",{}).get("e2e",{}).get("value"), "batched", d.get("batched",{}).get("value"))
PY
