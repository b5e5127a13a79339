timeout 900 python bench.py > gpurun_out/r02_bench_ours.json 2> gpurun_out/r02_bench_ours.err; tail -c 300 gpurun_out/r02_bench_ours.err
grep -o '"value": [0-9.]*, "unit": "nodes/s", "n_gpus": 1, "steps": 3, "warmup": 3, "ms_per_step": [0-9.]*' gpurun_out/r02_bench_ours.json
timeout 600 python -m pytest tests -m gpu -x -q -k "full_size" 2>&1 | tail -2
