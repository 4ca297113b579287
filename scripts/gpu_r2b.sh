#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2b_pytest.log
grep -E "passed|failed" gpurun_out/r2b_pytest.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/r2b_pytest.log | head -20
timeout 300 python bench.py --precision exact --steps 10 --warmup 3 --no-cpu > gpurun_out/r2b_bench_encodec_exact.json 2> gpurun_out/r2b_bench_encodec_exact.err
cat gpurun_out/r2b_bench_encodec_exact.json | head -c 600
