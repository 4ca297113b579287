#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r2v}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest.log
for spec in "encodec exact 64" "encodec fp16 64"; do
  set -- $spec
  AC_PRECISION=$2 timeout 400 python scripts/layer_times.py $1 $3 10 > gpurun_out/${T}_layers_$1_$2.txt 2>&1
  echo "$(grep '^total' gpurun_out/${T}_layers_$1_$2.txt || tail -2 gpurun_out/${T}_layers_$1_$2.txt)"
done
