mkdir -p gpurun_out
for d in 0 1 2 3 4 7; do ESVIO_K1_DBG=$d python scratch/stage_times.py stereo_vga_5mevs 40; done 2>&1 | tee gpurun_out/j5_variants.txt
python scratch/stage_times.py stereo_davis346_1mevs 40 2>&1 | tee -a gpurun_out/j5_variants.txt
python scratch/stage_times.py stereo_vga_20mevs_burst 30 2>&1 | tee -a gpurun_out/j5_variants.txt
