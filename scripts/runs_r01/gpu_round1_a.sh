set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc; free -g | head -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 1500 gpurun_out/bench_ref.json
timeout 600 python bench.py --batch 8 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_b8.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c2_b8.csv python bench.py --batch 8 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/ncu_launches.log
timeout 600 python bench.py --batch 8 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_b8.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:kvc_fused -s 6 -c 2 -o gpurun_out/prof_c2_b8 -f python bench.py --batch 8 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
