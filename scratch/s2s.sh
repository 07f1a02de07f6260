#!/bin/bash
T=${1:-s2s}
python -m pytest tests -m gpu -x -q -k "select or teacher or end_to_end or pipeline_equals or group_equals or lk_" > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
ESVIO_FE_LIB=$PWD/scratch/variants/libesvio_fe_clk.so python scratch/sel_clocks.py 2>&1 | grep window
python scratch/sync_timeline.py 2>&1 | grep -v Traceback | head -5
python bench.py --steps 40 --warmup 6 --no-cpu --no-frames --no-secondary --no-rigid > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f sync_ms %.3f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["sync_call_ms_per_step"], d["ms_per_step"]))
print("stage_ms", {k: round(v*1e3,1) for k,v in d["stage_ms"].items()})
r=d["roofline"]; print("roofline", r["frac"], r["kernel_ms"], r.get("frac_alone"), "batched", d.get("batched",{}).get("value"))
PY
