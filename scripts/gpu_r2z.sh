#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r2z}
for fm in 128 256; do
for prec in fp16 exact; do
  AC_FUSED_MAX_CH=$fm AC_PRECISION=$prec timeout 400 python scripts/layer_times.py encodec 64 10 > gpurun_out/${T}_layers_encodec_${prec}_fused$fm.txt 2>&1
  echo "fused<=$fm $(grep '^total' gpurun_out/${T}_layers_encodec_${prec}_fused$fm.txt || tail -3 gpurun_out/${T}_layers_encodec_${prec}_fused$fm.txt)"
done; done
grep "^times" gpurun_out/${T}_layers_encodec_fp16_fused256.txt | cut -c1-400
grep "^times" gpurun_out/${T}_layers_encodec_exact_fused256.txt | cut -c1-400
