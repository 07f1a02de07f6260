mkdir -p gpurun_out
T=j15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/${T}_bench_davis.json 2> gpurun_out/${T}_bench_davis.err
timeout 300 python bench.py --steps 200 --warmup 10 --workload stereo_vga_5mevs --cpu-windows 40 > gpurun_out/${T}_bench_vga.json 2> gpurun_out/${T}_bench_vga.err
timeout 300 python bench.py --impl reference --steps 60 --warmup 5 > gpurun_out/${T}_ref_davis.json 2> gpurun_out/${T}_ref_davis.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sae_update_ts -s 4 -c 1 -f -o gpurun_out/${T}_k1_davis python scratch/prof_k1.py stereo_davis346_1mevs 6 > gpurun_out/${T}_ncu_davis.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sae_update_ts -s 4 -c 1 -f -o gpurun_out/${T}_k1_vga python scratch/prof_k1.py stereo_vga_5mevs 6 > gpurun_out/${T}_ncu_vga.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/${T}_launches_davis.csv python bench.py --steps 12 --warmup 10 --no-cpu > gpurun_out/${T}_launch_bench.log 2>&1
timeout 300 python -m esvio_b200.replay --windows 31 > gpurun_out/${T}_replay.json 2>&1
python -c "
import json
for f in ('gpurun_out/${T}_bench_davis.json','gpurun_out/${T}_bench_vga.json'):
    d=json.load(open(f)); print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d.get('cpu_baseline',{}).get('value'), d.get('parity')); print(d['stage_ms'])
"
cat gpurun_out/${T}_replay.json
