#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r3a}
timeout 1500 python -m pytest tests -m gpu -q -x -rP > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest.log
grep -E "code match safe" gpurun_out/${T}_pytest.log | cut -c1-260
(time timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -4 gpurun_out/${T}_bench.err
python - <<P
import json
for l in open('gpurun_out/${T}_bench.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],'clocks',d['clocks'])
        print('parity',d.get('parity',{}).get('code_match_safe'),'fast',d.get('fast_mode',{}).get('value'),d.get('fast_mode',{}).get('parity',{}).get('code_match_safe'))
        for k,v in d.get('extra_configs',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('parity',{}).get('code_match_safe'), v.get('fast_mode',{}).get('value'), v.get('fast_mode',{}).get('parity',{}).get('code_match_safe'), v.get('error'))
P
