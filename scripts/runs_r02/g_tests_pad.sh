#!/bin/bash
# round 2, run g: full GPU suite (new full-size cases), streaming control with padded unit pitches, default bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log; tail -4 gpurun_out/r02g_pytest.log
for pad in 1 8 40 104; do
  python scripts/stream_copy_control.py --pad-rows $pad --only product,ldg_4cta --out gpurun_out/r02g_stream_pad.json 2>&1 | tail -2
done
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r02g_bench.err
