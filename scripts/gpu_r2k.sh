#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_resunit_tc_gpu.py tests/test_fp16_formats_gpu.py tests/test_exact_gpu.py tests/test_mimi_dac_gpu.py tests/test_dac_bf16_gpu.py -m gpu -q -x 2>&1 | tail -8
for io in 0 -1; do
  for c in dac encodec mimi; do
    B=64; [ $c = mimi ] && B=128
    AC_IO_STAGE=$io AC_PRECISION=fp16 timeout 400 python scripts/layer_times.py $c $B 10 > gpurun_out/r2k_layers_${c}_fp16_io$io.txt 2>&1
    echo "io=$io $(grep '^total' gpurun_out/r2k_layers_${c}_fp16_io$io.txt || tail -2 gpurun_out/r2k_layers_${c}_fp16_io$io.txt)"
  done
done
AC_IO_STAGE=0 AC_PRECISION=exact timeout 400 python scripts/layer_times.py encodec 64 10 | grep "^total"
AC_IO_STAGE=0 AC_PRECISION=exact timeout 400 python scripts/layer_times.py dac 64 10 | grep "^total"
