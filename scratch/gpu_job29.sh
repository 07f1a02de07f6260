mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 60 --warmup 10 > gpurun_out/j29_bench_n2.json 2> gpurun_out/j29_bench_n2.err
tail -c 1500 gpurun_out/j29_bench_n2.json; tail -5 gpurun_out/j29_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/j29_ref_n2.json 2> gpurun_out/j29_ref_n2.err
head -c 600 gpurun_out/j29_ref_n2.json
