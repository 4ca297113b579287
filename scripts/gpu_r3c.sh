#!/bin/bash
mkdir -p gpurun_out
for ew in 16 8; do
  echo "== AC_LSTM_EPI_WARPS=$ew"
  AC_LSTM_EPI_WARPS=$ew timeout 300 python -m pytest tests/test_encodec_bf16_gpu.py -m gpu -q -x -k lstm 2>&1 | tail -2
  AC_LSTM_EPI_WARPS=$ew timeout 120 python scripts/lstm_phase_profile.py 2>&1 | head -12
done
