DDO_FRINGE_PROF=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-config3 --no-cpu-baseline 2>&1 | grep -E '^\[solve\]' | cut -c1-600 | tail -2
timeout 300 python bench.py --steps 3 --warmup 2 --no-config3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*\|"host_fringe_ms_per_step": [0-9.]*\|"device_ms_per_step": [0-9.]*'
timeout 900 python -m pytest tests -m gpu -x -q -k "trajectory or sharded or solver or fringe or max2sat" 2>&1 | tail -2
