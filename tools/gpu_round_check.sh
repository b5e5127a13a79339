# Round-end validation on one B200 (run through gpurun): smoke, the default bench line, the reference arm, the whole GPU suite.
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_round_check.sh'   (outputs under gpurun_out/)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 300 gpurun_out/bench_ours.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.txt 2>&1; tail -2 gpurun_out/gpu_tests.txt
DDO_FRINGE_PROF=1 timeout 300 python bench.py --steps 2 --warmup 2 --no-config3 --no-cpu-baseline 2>&1 | grep -E '^\[solve\]' | cut -c1-600 | head -2 | tail -1 > gpurun_out/solve_phases.txt; cat gpurun_out/solve_phases.txt
head -c 400 gpurun_out/bench_ours.json
