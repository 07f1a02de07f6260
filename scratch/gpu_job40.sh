mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 445 -c 270 --csv --log-file gpurun_out/j40_launches_davis.csv python bench.py --steps 12 --warmup 10 --no-cpu --batch-streams 1 > gpurun_out/j40_launch_bench.log 2>&1
