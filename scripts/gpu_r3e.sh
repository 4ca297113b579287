#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/r3e_pytest.log | tail -1
for spec in "encodec exact 64" "encodec fp16 64" "encodec32 exact 64"; do
  set -- $spec
  AC_PRECISION=$2 timeout 400 python scripts/layer_times.py $1 $3 10 > gpurun_out/r3e_layers_$1_$2.txt 2>&1
  echo "$(grep '^total' gpurun_out/r3e_layers_$1_$2.txt || tail -2 gpurun_out/r3e_layers_$1_$2.txt)"; grep lstm_tc_kernel gpurun_out/r3e_layers_$1_$2.txt | cut -c1-40
done
timeout 300 python scripts/batch_sweep.py 2>&1 | tail -12
