mkdir -p gpurun_out
T=j18
timeout 400 python bench.py --steps 100 --warmup 10 --cpu-windows 20 > gpurun_out/${T}_bench_davis.json 2> gpurun_out/${T}_bench_davis.err
timeout 400 python bench.py --steps 100 --warmup 10 --workload stereo_vga_5mevs --cpu-windows 12 --batch-streams 4 > gpurun_out/${T}_bench_vga.json 2> gpurun_out/${T}_bench_vga.err
timeout 400 python bench.py --steps 60 --warmup 10 --workload stereo_vga_10mevs --no-cpu --batch-streams 4 > gpurun_out/${T}_bench_vga10.json 2> gpurun_out/${T}_bench_vga10.err
tail -3 gpurun_out/${T}_bench_davis.err gpurun_out/${T}_bench_vga.err gpurun_out/${T}_bench_vga10.err
python -c "
import json
for f in ('gpurun_out/${T}_bench_davis.json','gpurun_out/${T}_bench_vga.json','gpurun_out/${T}_bench_vga10.json'):
    try:
        d=json.load(open(f)); print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac']); print(' batched', json.dumps(d.get('batched')))
    except Exception as e: print(f, 'ERR', e)
"
