mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_select' -s 12 -c 1 -f -o gpurun_out/j35_select python bench.py --steps 30 --warmup 10 --no-cpu --batch-streams 1 > gpurun_out/j35_ncu.log 2>&1
tail -2 gpurun_out/j35_ncu.log
