#!/bin/bash
mkdir -p gpurun_out
AC_PRECISION=exact timeout 500 python scripts/encoder_error_probe.py 2 4 2>&1 | tail -4
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/r2g_pytest.log | tail -2; grep -E "^(FAILED|ERROR)" gpurun_out/r2g_pytest.log | head
grep -n "first differing" gpurun_out/r2g_pytest.log | cut -c1-400 | head -30
(time timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -3 gpurun_out/r2g_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
