#!/bin/bash
for st in 10 50; do
timeout 300 python bench.py --steps $st --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('steps $st: ms/step',d['ms_per_step'],'clocks',d['clocks'])
"
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --precision fp16 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('fp16: ms/step',d['ms_per_step'],'clocks',d['clocks'])
"
nvidia-smi --query-gpu=power.limit,power.default_limit,power.max_limit,clocks.max.sm --format=csv
