python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_2gpu_final.json 2> gpurun_out/r02_bench_2gpu_final.err
tail -c 200 gpurun_out/r02_bench_2gpu_final.err; head -c 1200 gpurun_out/r02_bench_2gpu_final.json
