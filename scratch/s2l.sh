#!/bin/bash
T=${1:-s2l}
python -m pytest tests -m gpu -x -q -k "good_features or track_image or frames" > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
python scratch/host_cost.py stereo_vga_5mevs 2>&1 | tail -4
python - <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
import bench
print(bench.frames_leg(640, 480, 0))
PY
