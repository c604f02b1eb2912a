#!/bin/bash
# round 2, run q (4 GPUs): the driver's launch line at N=4 (round 1's collapse point: 1 353 ms per e2e step)
mkdir -p gpurun_out
free -g | head -2
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 5 ) > gpurun_out/r02q_bench_n4.json 2> gpurun_out/r02q_bench_n4.err; echo "bench rc=$?"; tail -5 gpurun_out/r02q_bench_n4.err
