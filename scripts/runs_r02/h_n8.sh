#!/bin/bash
# round 2, run h (8 GPUs): the driver's launch line at N=8
mkdir -p gpurun_out
free -g | head -2
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r02h_bench_n8.json 2> gpurun_out/r02h_bench_n8.err; echo "bench rc=$?"; tail -5 gpurun_out/r02h_bench_n8.err
