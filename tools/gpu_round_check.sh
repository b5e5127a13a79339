python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_ours.json 2> gpurun_out/r02_bench_ours.err; tail -c 300 gpurun_out/r02_bench_ours.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.txt 2>&1; tail -3 gpurun_out/r02_gpu_tests.txt
head -c 400 gpurun_out/r02_bench_ours.json
