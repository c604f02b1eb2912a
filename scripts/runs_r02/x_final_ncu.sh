#!/bin/bash
# round 2, run x: ncu evidence with the final library — launch list of the bench command, full capture of the two c2
# launches, full capture of the fused vote launch at the c4 shape (B = 4)
mkdir -p gpurun_out
N="--no-e2e --no-cpu-baseline --no-configs --no-strong --no-eager --steps 2 --warmup 3"
python bench.py $N > gpurun_out/r02x_bench_plain.json 2> gpurun_out/r02x_bench_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kvc_ -c 400 --csv --log-file gpurun_out/r02x_launches_c2.csv python bench.py $N > gpurun_out/r02x_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kvc_fused -s 6 -c 2 -o gpurun_out/r02x_prof_c2 -f python bench.py $N > gpurun_out/r02x_ncu_full.log 2>&1; tail -1 gpurun_out/r02x_ncu_full.log
python bench.py --config c4_vote --batch 4 $N > /dev/null 2>&1; echo "vote plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kvc_snapkv -s 3 -c 1 -o gpurun_out/r02x_prof_vote -f python bench.py --config c4_vote --batch 4 $N > gpurun_out/r02x_ncu_vote.log 2>&1; tail -1 gpurun_out/r02x_ncu_vote.log
