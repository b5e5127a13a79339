timeout 300 python bench.py --steps 3 --warmup 2 --no-config3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*\|"host_fringe_ms_per_step": [0-9.]*\|"device_ms_per_step": [0-9.]*\|"kernel_ms_per_step": {[^}]*}'
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
