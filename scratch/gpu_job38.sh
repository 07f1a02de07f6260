mkdir -p gpurun_out
T=j38
timeout 400 python bench.py --steps 200 --warmup 10 --cpu-windows 30 > gpurun_out/${T}_bench_davis.json 2> gpurun_out/${T}_bench_davis.err
timeout 400 python bench.py --steps 200 --warmup 10 --workload stereo_vga_5mevs --cpu-windows 12 --batch-streams 4 > gpurun_out/${T}_bench_vga.json 2> gpurun_out/${T}_bench_vga.err
python -c "
import json
for f in ('gpurun_out/${T}_bench_davis.json','gpurun_out/${T}_bench_vga.json'):
    try:
        d=json.load(open(f)); print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']); print(d['stage_ms']); b=d.get('batched'); print(' batched', b['streams'], b['value'], b['ms_per_step'], b['roofline']['kernel_ms'], b['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
"
