#!/bin/bash
# 8-GPU runs of the final state: bench.py as the driver launches it (N = 8, also --codec dac) and the sharded K=32 sweep
mkdir -p gpurun_out
N=${N:-8}
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3) > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -3 gpurun_out/r02_bench_n$N.err
grep '^{' gpurun_out/r02_bench_n$N.json | head -c 700; echo
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --codec dac --no-extras) > gpurun_out/r02_bench_dac_n$N.json 2> gpurun_out/r02_bench_dac_n$N.err; grep '^{' gpurun_out/r02_bench_dac_n$N.json | head -c 400; echo
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/sharded_sweep.py 512 32) > gpurun_out/r02_sharded_sweep_n$N.jsonl 2> gpurun_out/r02_sharded_sweep_n$N.err; grep '^{' gpurun_out/r02_sharded_sweep_n$N.jsonl | tail -4
