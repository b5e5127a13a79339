DDO_FRINGE_PROF=1 timeout 300 python bench.py --steps 2 --warmup 2 --no-config3 --no-cpu-baseline 2>&1 | grep -E '^\[solve\]' | cut -c1-600 | head -2 | tail -1 > gpurun_out/r02_solve_phases.txt; cat gpurun_out/r02_solve_phases.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-config3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*\|"device_ms_per_step": [0-9.]*'
timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -2
