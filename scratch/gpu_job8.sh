mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,smsp__inst_executed.sum,launch__grid_size --clock-control none -k regex:'k_bin|k_pyr|k_corner' -s 24 -c 14 --csv --log-file gpurun_out/j8_k0.csv python scratch/prof_k1.py stereo_vga_5mevs 6 > gpurun_out/j8.log 2>&1
