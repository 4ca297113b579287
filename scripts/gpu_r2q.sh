#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_encodec_bf16_gpu.py -m gpu -q -x -k lstm 2>&1 | tail -3
timeout 120 python scripts/lstm_phase_profile.py 2>&1
