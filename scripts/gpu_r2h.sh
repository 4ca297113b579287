#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_consumers.py -m gpu -q 2>&1 | tail -3
timeout 900 python scripts/profile_protocol.py encodec dac mimi > gpurun_out/r2h_profile_protocol.jsonl 2> gpurun_out/r2h_profile_protocol.err; tail -2 gpurun_out/r2h_profile_protocol.err; cut -c1-420 gpurun_out/r2h_profile_protocol.jsonl
