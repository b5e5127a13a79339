python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_bench_ours.json 2> gpurun_out/r02_bench_ours.err; tail -c 300 gpurun_out/r02_bench_ours.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.txt 2>&1; tail -2 gpurun_out/r02_gpu_tests.txt
DDO_FRINGE_PROF=1 timeout 300 python bench.py --steps 2 --warmup 2 --no-config3 --no-cpu-baseline 2>&1 | grep -E '^\[solve\]' | cut -c1-600 | head -2 | tail -1 > gpurun_out/r02_solve_phases.txt; cat gpurun_out/r02_solve_phases.txt
grep -o '"value": [0-9.]*, "unit": "nodes/s", "n_gpus": 1, "steps": 3, "warmup": 3, "ms_per_step": [0-9.]*' gpurun_out/r02_bench_ours.json
