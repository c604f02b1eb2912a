set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2_v2.json 2> gpurun_out/bench_c2_v2.err; tail -c 4000 gpurun_out/bench_c2_v2.json; tail -5 gpurun_out/bench_c2_v2.err
