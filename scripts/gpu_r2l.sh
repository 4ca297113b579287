#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/r2l_pytest.log | tail -2; grep -E "^(FAILED|ERROR)" gpurun_out/r2l_pytest.log | head
for c in dac encodec mimi; do
  B=64; [ $c = mimi ] && B=128
  for prec in fp16 exact; do
  AC_PRECISION=$prec timeout 400 python scripts/layer_times.py $c $B 10 > gpurun_out/r2l_layers_${c}_$prec.txt 2>&1
  echo "$(grep '^total' gpurun_out/r2l_layers_${c}_$prec.txt || tail -2 gpurun_out/r2l_layers_${c}_$prec.txt)"
  done
done
grep -h "^tuned" gpurun_out/r2l_layers_*_fp16.txt | head -30
