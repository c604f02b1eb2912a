#!/bin/bash
# round 2, run s: final library after the re-entry session — full GPU suite, smoke, both bench arms
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02s_pytest.log; tail -3 gpurun_out/r02s_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02s_bench_ref.json 2> gpurun_out/r02s_bench_ref.err; echo "ref rc=$?"
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r02s_bench.err
