mkdir -p gpurun_out
T=j47
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench_davis.json 2> gpurun_out/${T}_bench_davis.err
timeout 600 python bench.py --workload stereo_vga_5mevs --batch-streams 4 > gpurun_out/${T}_bench_vga.json 2> gpurun_out/${T}_bench_vga.err
timeout 600 python bench.py --workload stereo_vga_10mevs --batch-streams 8 --no-cpu --steps 60 > gpurun_out/${T}_bench_vga10.json 2> gpurun_out/${T}_bench_vga10.err
python -c "
import json
for f in ('gpurun_out/${T}_bench_davis.json','gpurun_out/${T}_bench_vga.json','gpurun_out/${T}_bench_vga10.json'):
    try:
        d=json.load(open(f)); r=d['roofline']; print(f, d['value'], d['e2e']['value'], d['ms_per_step'], r['frac'], r['frac_alone'], d.get('cpu_baseline',{}).get('value')); b=d.get('batched'); print(' batched', b['streams'], b['value'], b['ms_per_step'], b['roofline']['kernel_ms'], b['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
"
