#!/bin/bash
T=${1:-s2v}
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_bench.json 2> gpurun_out/${T}_ncu_bench.err
python scratch/launch_summary.py gpurun_out/${T}_launches.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 python bench.py --steps 8 --warmup 3 --no-cpu (cold-cache, serialised: compare shares)" | tee gpurun_out/${T}_launches.txt
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 500 --csv --log-file gpurun_out/${T}_launches1.csv python bench.py --steps 16 --warmup 4 --no-cpu --no-frames --no-secondary --no-rigid --batch-streams 1 > /dev/null 2>&1
python scratch/launch_summary.py gpurun_out/${T}_launches1.csv "single stream stereo_vga_5mevs only (--no-frames --no-secondary --no-rigid --batch-streams 1), launches 200..700" | tee gpurun_out/${T}_launches1.txt
