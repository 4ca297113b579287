#!/bin/bash
# epilogue instruction diet (bare MUFU.EX2 ELU, LDS/STS instead of generic LD/ST): parity tests + per-layer times
mkdir -p gpurun_out
T=${TAG:-r2o}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
for spec in "encodec fp16 64" "encodec exact 64" "dac fp16 64" "mimi fp16 128" "dac exact 64" "mimi exact 128"; do
  set -- $spec
  AC_PRECISION=$2 timeout 400 python scripts/layer_times.py $1 $3 10 > gpurun_out/${T}_layers_$1_$2.txt 2>&1
  echo "$(grep '^total' gpurun_out/${T}_layers_$1_$2.txt || tail -2 gpurun_out/${T}_layers_$1_$2.txt)"
done
timeout 120 python scripts/lstm_cluster_occupancy.py > gpurun_out/${T}_lstm_occupancy.txt 2>&1; cat gpurun_out/${T}_lstm_occupancy.txt | head -20
