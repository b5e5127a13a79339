#!/bin/bash
# A/B runs of alternative builds of the library (development): kernel times of one bench pass per variant
for v in "$@"; do
  DDO_B200_LIB=ddo_b200/libddo_b200_$v.so python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ab_$v.json 2>&1
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],1), round(d['wall_ms_per_step'],1), d['roofline']['kernel_ms_per_step'])"
done
