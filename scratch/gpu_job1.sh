set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/j1_smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/j1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j1_pytest.log
timeout 300 python bench.py --steps 90 --warmup 10 --cpu-windows 40 > gpurun_out/j1_bench_davis.json 2> gpurun_out/j1_bench_davis.err
timeout 300 python bench.py --steps 90 --warmup 10 --workload stereo_vga_5mevs --cpu-windows 30 > gpurun_out/j1_bench_vga.json 2> gpurun_out/j1_bench_vga.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sae_update_ts -s 4 -c 2 -f -o gpurun_out/j1_k1_vga python scratch/prof_k1.py stereo_vga_5mevs 7 > gpurun_out/j1_ncu_vga.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sae_update_ts -s 4 -c 1 -f -o gpurun_out/j1_k1_davis python scratch/prof_k1.py stereo_davis346_1mevs 6 > gpurun_out/j1_ncu_davis.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/j1_launches_vga.csv python bench.py --steps 12 --warmup 10 --no-cpu --workload stereo_vga_5mevs > gpurun_out/j1_launch_bench.log 2>&1
tail -3 gpurun_out/j1_pytest.log
cat gpurun_out/j1_bench_davis.json | head -c 1500
