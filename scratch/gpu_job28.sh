mkdir -p gpurun_out
T=j28
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_pytest.log
python scratch/stage_times.py stereo_davis346_1mevs 40 2>&1 | tee gpurun_out/${T}_stage.txt
python scratch/stage_times.py stereo_vga_5mevs 40 2>&1 | tee -a gpurun_out/${T}_stage.txt
python scratch/group_k1.py stereo_davis346_1mevs 8 2>&1 | tee gpurun_out/${T}_group_k1.txt
python scratch/group_k1.py stereo_vga_5mevs 4 8 2>&1 | tee -a gpurun_out/${T}_group_k1.txt
python scratch/group_k1.py stereo_vga_10mevs 8 2>&1 | tee -a gpurun_out/${T}_group_k1.txt
