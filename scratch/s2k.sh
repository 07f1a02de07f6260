#!/bin/bash
T=${1:-s2k}
for co in 50 75 100; do
echo "== carveout $co"
ESVIO_CARVEOUT=$co python bench.py --steps 40 --warmup 6 --no-cpu --no-frames --no-secondary --no-rigid > gpurun_out/${T}_bench$co.json 2> gpurun_out/${T}_bench$co.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench$co.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f sync_ms %.3f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["sync_call_ms_per_step"], d["ms_per_step"]))
print("stage_ms", {k: round(v*1e3,1) for k,v in d["stage_ms"].items()})
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "batched", d.get("batched",{}).get("value"))
PY
done
