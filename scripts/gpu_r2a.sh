#!/bin/bash
# round 2, GPU call A: full GPU test suite + per-layer times / bench lines of the exact and bf16 precision modes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
for prec in exact bf16; do
  AC_PRECISION=$prec timeout 300 python scripts/layer_times.py encodec 64 10 > gpurun_out/r2a_layers_encodec_$prec.txt 2>&1
  timeout 300 python bench.py --precision $prec --steps 10 --warmup 3 --no-cpu > gpurun_out/r2a_bench_encodec_$prec.json 2> gpurun_out/r2a_bench_encodec_$prec.err
  tail -1 gpurun_out/r2a_layers_encodec_$prec.txt | head -c 300; echo
done
AC_PRECISION=exact AC_TUNE_FUSION=1 timeout 300 python scripts/layer_times.py encodec 64 10 > gpurun_out/r2a_layers_encodec_exact_tunefusion.txt 2>&1
for c in dac mimi; do
  B=64; [ $c = mimi ] && B=128
  for prec in exact bf16; do
    AC_PRECISION=$prec timeout 400 python scripts/layer_times.py $c $B 10 > gpurun_out/r2a_layers_${c}_$prec.txt 2>&1
    grep "^total" gpurun_out/r2a_layers_${c}_$prec.txt
  done
done
grep -h "^total" gpurun_out/r2a_layers_*.txt
