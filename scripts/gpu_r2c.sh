#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -3 gpurun_out/r2c_bench.err
(time timeout 300 python bench.py --impl reference --steps 5 --warmup 2) > gpurun_out/r2c_bench_reference.json 2> gpurun_out/r2c_bench_reference.err; tail -3 gpurun_out/r2c_bench_reference.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.txt 2>&1; tail -2 gpurun_out/r2c_smoke.txt
head -c 1500 gpurun_out/r2c_bench.json
