#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_encodec_bf16_gpu.py -m gpu -q -x -k lstm 2>&1 | tail -2
for cl in 7 4; do for tr in 0 1; do
  echo "== clusters $cl trim $tr"; AC_LSTM_CLUSTERS=$cl AC_LSTM_TRIM=$tr timeout 120 python scripts/lstm_phase_profile.py 2>&1 | head -12
done; done
