#!/bin/bash
T=${1:-s2h}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
for w in stereo_vga_5mevs stereo_davis346_1mevs; do python scratch/stage_times.py $w 40 2>&1 | tail -1; done
echo "== warm per-kernel times (ncu --cache-control none), sync windows"
ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum -s 200 -c 300 --csv --log-file gpurun_out/${T}_warm.csv python scratch/stage_times.py stereo_vga_5mevs 40 > /dev/null 2>&1
python scratch/launch_summary.py gpurun_out/${T}_warm.csv "warm, sync windows, vga5" | tee gpurun_out/${T}_warm.txt
python scratch/group_host_cost.py stereo_vga_5mevs 8 2>&1 | tail -1
