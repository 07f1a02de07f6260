#!/bin/bash
T=s2e
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
echo "== group K1 alone"; python scratch/group_k1.py stereo_vga_5mevs 1 8 2>&1 | tail -2
echo "== group K1 alone, K1 carveout 100"; ESVIO_CARVEOUT_K1=100 python scratch/group_k1.py stereo_vga_5mevs 1 8 2>&1 | tail -2
summ() { python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_$1.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f sync_ms %.3f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["sync_call_ms_per_step"], d["ms_per_step"]))
print("stage_ms", {k: round(v*1e3,1) for k,v in d["stage_ms"].items()})
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "batched", d.get("batched",{}).get("value"))
PY
}
echo "== bench default"; python bench.py --steps 30 --warmup 6 --no-cpu --no-frames --no-secondary --no-rigid > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; summ bench
echo "== bench K1 carveout 100"; ESVIO_CARVEOUT_K1=100 python bench.py --steps 30 --warmup 6 --no-cpu --no-frames --no-secondary --no-rigid > gpurun_out/${T}_benchk100.json 2> gpurun_out/${T}_benchk100.err; summ benchk100
echo "== launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 16 --warmup 4 --no-cpu --no-frames --no-secondary --no-rigid --batch-streams 1 > /dev/null 2>&1
python scratch/launch_summary.py gpurun_out/${T}_launches.csv "s2e single stream vga5" | tee gpurun_out/${T}_launches.txt
