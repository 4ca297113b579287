#!/bin/bash
# same-box A/B of the whole step: shipped library vs the same library with the one-group LSTM kernel (16 epilogue warps)
for rep in 1 2 3; do
for lib in new old; do
  P=""; [ $lib = old ] && P=/root/repo/audiocodecs_b200/lib/libaudiocodecs_b200_lstmv3.so
  AC_LIB_PATH=$P timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['all_kernels_ms']
        print('$lib rep $rep: ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'lstm',k['lstm_tc_kernel'],'conv_tc',k['conv_tc_kernel'],'resunit',k['resunit_tc_kernel'],'rvq',k['rvq_encode_tc_kernel'],'clocks',d['clocks']['sm_mhz'])
"
done; done
