#!/bin/bash
# round 2, run b: remaining GPU tests + the default bench line (e2e via stored norms, configs table, strong section)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -25 gpurun_out/r02b_pytest.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02b_bench.err
head -c 3000 gpurun_out/r02b_bench.json
