mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/j44_pytest.log
python scratch/stage_times.py stereo_davis346_1mevs 40 2>&1 | tee gpurun_out/j44_stage.txt
python scratch/stage_times.py stereo_vga_5mevs 40 2>&1 | tee -a gpurun_out/j44_stage.txt
