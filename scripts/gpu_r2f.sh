#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/accum_probe.py 2>&1 | grep "K= 4096\|K= 1024" 
AC_PRECISION=exact timeout 500 python scripts/encoder_error_probe.py 2 4 2>&1 | tail -15
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/r2f_pytest.log | tail -2; grep -E "^(FAILED|ERROR)" gpurun_out/r2f_pytest.log | head
grep -n "code match safe\|chunked\|rel-err" gpurun_out/r2f_pytest.log | grep -v "bf16 " | head -30
AC_PRECISION=exact timeout 300 python scripts/layer_times.py encodec 64 10 | grep "^total\|lstm_ih\|down_tc\|conv_k7\|res_k3\|res_tail"
