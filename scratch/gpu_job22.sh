mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_lk|k_select|k_ransac' -s 60 -c 8 -f -o gpurun_out/j22_track python bench.py --steps 12 --warmup 10 --no-cpu --batch-streams 1 > gpurun_out/j22_ncu.log 2>&1
tail -3 gpurun_out/j22_ncu.log
