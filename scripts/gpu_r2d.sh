#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2d_pytest.log
grep -E "passed|failed" gpurun_out/r2d_pytest.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/r2d_pytest.log | head -30
for c in encodec dac mimi; do
  B=64; [ $c = mimi ] && B=128
  for prec in exact fp16; do
    AC_PRECISION=$prec timeout 400 python scripts/layer_times.py $c $B 10 > gpurun_out/r2d_layers_${c}_$prec.txt 2>&1
    grep "^total" gpurun_out/r2d_layers_${c}_$prec.txt || tail -3 gpurun_out/r2d_layers_${c}_$prec.txt
  done
done
