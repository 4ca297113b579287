#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r2w}
timeout 600 python -m pytest tests/test_resunit_tc_gpu.py -m gpu -q -x 2>&1 | tail -5
for spec in "encodec fp16 64" "encodec exact 64" "dac fp16 64" "mimi fp16 128"; do
  set -- $spec
  AC_PRECISION=$2 timeout 400 python scripts/layer_times.py $1 $3 10 > gpurun_out/${T}_layers_$1_$2.txt 2>&1
  echo "$(grep '^total' gpurun_out/${T}_layers_$1_$2.txt || tail -2 gpurun_out/${T}_layers_$1_$2.txt)"
  grep "^tuned" gpurun_out/${T}_layers_$1_$2.txt | cut -c1-130
done
