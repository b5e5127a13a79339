timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_2gpu_final.json 2> gpurun_out/r02_bench_2gpu_final.err
tail -c 200 gpurun_out/r02_bench_2gpu_final.err; grep -o '"value": [0-9.]*, "unit": "nodes/s", "n_gpus": 2, "steps": 3, "warmup": 2, "ms_per_step": [0-9.]*\|"expanded_nodes_per_step": [0-9]*\|"device_ms_per_step": [0-9.]*' gpurun_out/r02_bench_2gpu_final.json
