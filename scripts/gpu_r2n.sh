#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_resunit_tc_gpu.py tests/test_conv_tc_gpu.py tests/test_fp16_formats_gpu.py tests/test_exact_gpu.py -m gpu -q -x 2>&1 | tail -3
for c in dac encodec mimi; do
  B=64; [ $c = mimi ] && B=128
  AC_PRECISION=fp16 timeout 400 python scripts/layer_times.py $c $B 10 > gpurun_out/r2n_layers_${c}_fp16.txt 2>&1
  echo "$(grep '^total' gpurun_out/r2n_layers_${c}_fp16.txt || tail -2 gpurun_out/r2n_layers_${c}_fp16.txt)"
done
AC_PRECISION=exact timeout 400 python scripts/layer_times.py encodec 64 10 | grep "^total"
