#!/bin/bash
# K1 experiments: TS fast path + L2 hints + carve-out
T=s2c
python -m pytest tests -m gpu -x -q -k "time_surface or sae or teacher or group or conditioning or ignore" > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
for pin in 0 6 9 12 16; do
  echo "== pin_cams=$pin"; ESVIO_K1_PIN_CAMS=$pin python scratch/group_k1.py stereo_vga_5mevs 1 8 2>&1 | tail -2
done
echo "== carveout 100, default pin"; ESVIO_CARVEOUT=100 python scratch/group_k1.py stereo_vga_5mevs 1 8 2>&1 | tail -2
echo "== bench default"; python bench.py --steps 30 --warmup 6 --no-cpu --no-frames --no-secondary --no-rigid > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f sync_ms %.3f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["sync_call_ms_per_step"], d["ms_per_step"]))
print("stage_ms", {k: round(v*1e3,1) for k,v in d["stage_ms"].items()})
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "batched", d.get("batched",{}).get("value"))
PY
echo "== bench carveout 100"; ESVIO_CARVEOUT=100 python bench.py --steps 30 --warmup 6 --no-cpu --no-frames --no-secondary --no-rigid > gpurun_out/${T}_bench100.json 2> gpurun_out/${T}_bench100.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench100.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f sync_ms %.3f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["sync_call_ms_per_step"], d["ms_per_step"]))
print("stage_ms", {k: round(v*1e3,1) for k,v in d["stage_ms"].items()})
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "batched", d.get("batched",{}).get("value"))
PY
