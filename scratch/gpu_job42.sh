mkdir -p gpurun_out
T=j42
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_pytest.log
timeout 600 python bench.py --batch-streams 1 > gpurun_out/${T}_bench_davis.json 2> gpurun_out/${T}_bench_davis.err
timeout 600 python bench.py --workload stereo_vga_5mevs --batch-streams 1 --cpu-windows 12 > gpurun_out/${T}_bench_vga.json 2> gpurun_out/${T}_bench_vga.err
python -c "
import json
for f in ('gpurun_out/${T}_bench_davis.json','gpurun_out/${T}_bench_vga.json'):
    try:
        d=json.load(open(f)); r=d['roofline']; print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['sync_call_ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
"
