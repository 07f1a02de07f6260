python scratch/group_k1.py stereo_davis346_1mevs 1 2 4 8 2>&1 | tee gpurun_out/j21_group_k1.txt
python scratch/group_k1.py stereo_vga_5mevs 1 2 4 8 2>&1 | tee -a gpurun_out/j21_group_k1.txt
python scratch/group_k1.py stereo_vga_10mevs 4 8 2>&1 | tee -a gpurun_out/j21_group_k1.txt
