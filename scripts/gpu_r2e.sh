#!/bin/bash
mkdir -p gpurun_out
AC_PRECISION=exact timeout 600 python scripts/encoder_error_probe.py 2 4 > gpurun_out/r2e_probe_exact.txt 2>&1; cat gpurun_out/r2e_probe_exact.txt | tail -20
timeout 900 python -m pytest tests/test_full_size_gpu.py tests/test_mimi_bf16_gpu.py tests/test_fp16_formats_gpu.py -m gpu -q 2>&1 | tail -5
