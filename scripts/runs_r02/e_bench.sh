#!/bin/bash
# round 2, run e: e2e probe (where the host-resident step spends its time) + the full default bench line
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "compiled_binding or current_device" > gpurun_out/r02e_pytest.log 2>&1; tail -3 gpurun_out/r02e_pytest.log
python scripts/e2e_probe.py --out gpurun_out/r02e_e2e_probe.json > gpurun_out/r02e_e2e_probe.log 2>&1; cat gpurun_out/r02e_e2e_probe.log | tail -8
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r02e_bench.err
