#!/bin/bash
T=${1:-s2m}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f sync_ms %.3f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["sync_call_ms_per_step"], d["ms_per_step"]))
print("stage_ms", {k: round(v*1e3,1) for k,v in d["stage_ms"].items()})
r=d["roofline"]; print("roofline", r["frac"], r["kernel_ms"], r.get("frac_alone"), r.get("kernel_ms_alone"), "batched", d.get("batched",{}).get("value"))
print("rigid", d["rigid_scene"]["value"], d["rigid_scene"]["e2e"]["value"], "frames", d["frames"].get("value"), "cpu", d["cpu_baseline"]["value"])
PY
