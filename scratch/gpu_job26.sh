USE_RANSAC=0 python scratch/stage_times.py stereo_davis346_1mevs 40 2>&1 | tee gpurun_out/j26_stage.txt
USE_RANSAC=0 python scratch/stage_times.py stereo_vga_5mevs 40 2>&1 | tee -a gpurun_out/j26_stage.txt
