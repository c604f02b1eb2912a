set -x
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
$B --config c2 --steps 2 > gpurun_out/plain_c2.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_final.csv -k regex:kvc_ $B --config c2 --steps 2 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
$B --config c2 --steps 1 > gpurun_out/plain_c2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:kvc_fused -s 6 -c 2 -o gpurun_out/prof_c2_final -f $B --config c2 --steps 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
