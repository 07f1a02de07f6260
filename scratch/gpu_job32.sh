mkdir -p gpurun_out
T=j32
timeout 600 python bench.py > gpurun_out/${T}_bench_davis.json 2> gpurun_out/${T}_bench_davis.err
timeout 600 python bench.py --workload stereo_vga_5mevs --batch-streams 4 > gpurun_out/${T}_bench_vga.json 2> gpurun_out/${T}_bench_vga.err
timeout 600 python bench.py --workload stereo_vga_10mevs --batch-streams 8 --no-cpu --steps 60 > gpurun_out/${T}_bench_vga10.json 2> gpurun_out/${T}_bench_vga10.err
timeout 300 python bench.py --impl reference --steps 60 --warmup 5 > gpurun_out/${T}_ref_davis.json 2> gpurun_out/${T}_ref_davis.err
timeout 300 python bench.py --impl reference --steps 30 --warmup 5 --workload stereo_vga_5mevs > gpurun_out/${T}_ref_vga.json 2> gpurun_out/${T}_ref_vga.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/${T}_launches_davis.csv python bench.py --steps 12 --warmup 10 --no-cpu --batch-streams 1 > gpurun_out/${T}_launch_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sae_update_ts -s 4 -c 1 -f -o gpurun_out/${T}_k1_davis python scratch/prof_k1.py stereo_davis346_1mevs 6 > gpurun_out/${T}_ncu_davis.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sae_update_ts -s 4 -c 1 -f -o gpurun_out/${T}_k1_vga python scratch/prof_k1.py stereo_vga_5mevs 6 > gpurun_out/${T}_ncu_vga.log 2>&1
python scratch/group_k1.py stereo_davis346_1mevs 8 2>&1 | tee gpurun_out/${T}_group_k1.txt
python scratch/group_k1.py stereo_vga_5mevs 4 8 2>&1 | tee -a gpurun_out/${T}_group_k1.txt
python scratch/group_k1.py stereo_vga_10mevs 4 8 2>&1 | tee -a gpurun_out/${T}_group_k1.txt
python -c "
import json
for f in ('gpurun_out/${T}_bench_davis.json','gpurun_out/${T}_bench_vga.json','gpurun_out/${T}_bench_vga10.json','gpurun_out/${T}_ref_davis.json','gpurun_out/${T}_ref_vga.json'):
    try:
        d=json.load(open(f)); print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d.get('roofline',{}).get('frac'), d.get('cpu_baseline',{}).get('value'), d.get('parity')); b=d.get('batched'); 
        if b: print(' batched', b['streams'], b['value'], b['ms_per_step'], b['roofline']['kernel_ms'], b['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
"
