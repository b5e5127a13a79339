#!/bin/bash
# A/B runs of one bench pass under different environment settings: tools/ab_env.sh "VAR=1 VAR2=3" "VAR=2" ...
i=0
for e in "$@"; do
  i=$((i+1))
  env $e python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/abenv_$i.json 2>&1
  python -c "
import json; d=json.loads(open('gpurun_out/abenv_$i.json').read().strip().splitlines()[-1]); print('$e', round(d['ms_per_step'],1), round(d['wall_ms_per_step'],1), d['roofline']['kernel_ms_per_step'])"
done
