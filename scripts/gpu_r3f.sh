#!/bin/bash
for tr in 3 1; do
  echo "== trim $tr"; AC_LSTM_TRIM=$tr timeout 300 python -m pytest tests/test_encodec_bf16_gpu.py -m gpu -q -x -k lstm 2>&1 | tail -1
  AC_LSTM_TRIM=$tr timeout 120 python scripts/lstm_phase_profile.py 2>&1 | head -12
done
