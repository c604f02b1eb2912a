#!/bin/bash
# round 2, run f (2 GPUs): the driver's launch line at N=2 — weak-scaled c2, e2e per rank, strong section; 2-GPU device test
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py -m gpu -q -x -k "current_device or pinned_slab" > gpurun_out/r02f_pytest.log 2>&1; tail -3 gpurun_out/r02f_pytest.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err; echo "bench rc=$?"; tail -5 gpurun_out/r02f_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/r02f_bench_ref_n2.json 2>/dev/null; echo "ref rc=$?"; head -c 300 gpurun_out/r02f_bench_ref_n2.json
