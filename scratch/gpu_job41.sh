mkdir -p gpurun_out
T=j41
timeout 600 python bench.py > gpurun_out/${T}_bench_davis.json 2> gpurun_out/${T}_bench_davis.err
timeout 600 python bench.py --workload stereo_vga_5mevs --batch-streams 4 > gpurun_out/${T}_bench_vga.json 2> gpurun_out/${T}_bench_vga.err
tail -3 gpurun_out/${T}_bench_davis.err
python -c "
import json
for f in ('gpurun_out/${T}_bench_davis.json','gpurun_out/${T}_bench_vga.json'):
    try:
        d=json.load(open(f)); r=d['roofline']; print(f, d['value'], d['e2e']['value'], d['ms_per_step'], r['frac'], r['frac_alone'], r['kernel_ms'], r['kernel_ms_alone'], d['gpu_launches']); print(d['stage_ms'])
    except Exception as e: print(f, 'ERR', e)
"
