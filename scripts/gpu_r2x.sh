#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r2x}
for raw in 32 64; do
for prec in fp16 exact; do
  AC_RAW_MAX_CH=$raw AC_PRECISION=$prec timeout 400 python scripts/layer_times.py encodec 64 10 > gpurun_out/${T}_layers_encodec_${prec}_raw$raw.txt 2>&1
  echo "raw<=$raw $(grep '^total' gpurun_out/${T}_layers_encodec_${prec}_raw$raw.txt || tail -3 gpurun_out/${T}_layers_encodec_${prec}_raw$raw.txt)"
done; done
AC_RAW_MAX_CH=64 timeout 600 python -m pytest tests/test_exact_gpu.py tests/test_encodec_bf16_gpu.py tests/test_resunit_tc_gpu.py -m gpu -q -x 2>&1 | tail -4
