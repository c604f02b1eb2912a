#!/bin/bash
# round 2, run d: compiled binding — tests, per-call host overhead at batch 1, the decode-loop harness, a short bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -6 gpurun_out/r02d_pytest.log
python scripts/host_overhead.py > gpurun_out/r02d_host_overhead.log 2>&1; head -14 gpurun_out/r02d_host_overhead.log
python scripts/harness_bench.py > gpurun_out/r02d_harness.json 2> gpurun_out/r02d_harness.err; echo "harness rc=$?"; tail -c 600 gpurun_out/r02d_harness.err
python bench.py --no-e2e --no-cpu-baseline --no-strong --no-eager --steps 10 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"
