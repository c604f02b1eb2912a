#!/bin/bash
# round 2, run n: final library — full GPU suite, smoke, both bench arms, launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest.log; tail -3 gpurun_out/r02n_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02n_bench_ref.json 2> gpurun_out/r02n_bench_ref.err; echo "ref rc=$?"
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r02n_bench.err
N="--no-e2e --no-cpu-baseline --no-configs --no-strong --no-eager --steps 2 --warmup 3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kvc_ -c 400 --csv --log-file gpurun_out/r02n_launches_c2.csv python bench.py $N > gpurun_out/r02n_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kvc_fused -s 6 -c 2 -o gpurun_out/r02n_prof_c2 -f python bench.py $N > gpurun_out/r02n_ncu_full.log 2>&1; tail -1 gpurun_out/r02n_ncu_full.log
