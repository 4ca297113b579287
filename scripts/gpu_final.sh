#!/bin/bash
# Final-state evidence of the round: full GPU test suite, smoke(), bench line, per-layer times, ncu launch lists.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/r02_pytest_gpu.log | tail -1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke.log
(time timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -4 gpurun_out/r02_bench.err
(timeout 600 python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; head -c 600 gpurun_out/r02_bench_reference.json; echo
for spec in "encodec exact 64" "encodec fp16 64" "dac exact 64" "dac fp16 64" "mimi exact 128" "mimi fp16 128"; do
  set -- $spec
  AC_PRECISION=$2 timeout 400 python scripts/layer_times.py $1 $3 10 > gpurun_out/r02_layers_$1_$2.txt 2>&1
  echo "$(grep '^total' gpurun_out/r02_layers_$1_$2.txt || tail -2 gpurun_out/r02_layers_$1_$2.txt)"
done
bash scripts/gpu_ncu_lists.sh > gpurun_out/r02_ncu_lists.log 2>&1; tail -3 gpurun_out/r02_ncu_lists.log
