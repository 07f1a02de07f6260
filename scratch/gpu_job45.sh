mkdir -p gpurun_out
T=j45
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
tail -2 gpurun_out/${T}_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29528 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/${T}_ref_n2.json 2> gpurun_out/${T}_ref_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/j45_bench_n2.json','gpurun_out/j45_ref_n2.json'):
    lines=[l for l in open(f) if l.strip()]
    print(f, len(lines), 'line(s)')
    for l in lines:
        try:
            d=json.loads(l); print(' ', d.get('impl','ours'), d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'])
        except Exception as e: print('  NONJSON', l[:120])
PY
