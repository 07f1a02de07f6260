mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_select' -s 3 -c 2 -f -o gpurun_out/j25_select python bench.py --steps 12 --warmup 10 --no-cpu --batch-streams 1 > gpurun_out/j25_ncu.log 2>&1
tail -2 gpurun_out/j25_ncu.log
