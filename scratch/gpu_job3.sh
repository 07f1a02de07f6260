set -x
mkdir -p gpurun_out
T=j4
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 90 --warmup 10 --cpu-windows 12 > gpurun_out/${T}_bench_davis.json 2> gpurun_out/${T}_bench_davis.err
timeout 300 python bench.py --steps 90 --warmup 10 --workload stereo_vga_5mevs --cpu-windows 8 > gpurun_out/${T}_bench_vga.json 2> gpurun_out/${T}_bench_vga.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_bin|k_sae|k_corner|k_pyr' -s 32 -c 8 -f -o gpurun_out/${T}_evstage_vga python scratch/prof_k1.py stereo_vga_5mevs 6 > gpurun_out/${T}_ncu_vga.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/${T}_launches_vga.csv python bench.py --steps 12 --warmup 10 --no-cpu --workload stereo_vga_5mevs > gpurun_out/${T}_launch_bench.log 2>&1
python -c "
import json
for f in ('gpurun_out/${T}_bench_davis.json','gpurun_out/${T}_bench_vga.json'):
    d=json.load(open(f)); print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac']); print(d['stage_ms'])
"
