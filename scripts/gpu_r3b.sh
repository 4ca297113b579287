#!/bin/bash
mkdir -p gpurun_out
for pair in 1 0; do
  echo "== AC_RVQ_PAIR=$pair"
  AC_RVQ_PAIR=$pair timeout 600 python -m pytest tests/test_encodec_gpu.py tests/test_exact_gpu.py -m gpu -q -x -k "rvq or exact or K32 or full" 2>&1 | tail -2
  AC_RVQ_PAIR=$pair AC_PRECISION=fp16 timeout 300 python scripts/layer_times.py encodec 64 10 2>&1 | grep -E "rvq_encode|^total"
  AC_RVQ_PAIR=$pair AC_PRECISION=fp16 timeout 300 python scripts/layer_times.py encodec32 64 10 2>&1 | grep -E "rvq_encode|^total"
done
