for d in 0 8; do ESVIO_K1_DBG=$d python scratch/stage_times.py stereo_vga_5mevs 40; done 2>&1 | tee gpurun_out/j10_variants.txt
