#!/bin/bash
timeout 300 python -m pytest tests/test_encodec_bf16_gpu.py -m gpu -q -x -k lstm 2>&1 | tail -1
for rep in 1 2; do
for cfg in "7 0" "7 1" "4 0"; do
  set -- $cfg
  AC_LSTM_CLUSTERS=$1 AC_LSTM_POLL=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['all_kernels_ms']
        print('clusters $1 poll $2 rep $rep: ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'lstm',k['lstm_tc_kernel'],'conv_tc',k['conv_tc_kernel'],'resunit',k['resunit_tc_kernel'],'clocks',d['clocks']['sm_mhz'])
"
done; done
AC_LSTM_POLL=0 timeout 120 python scripts/lstm_phase_profile.py 2>&1 | head -3
