#!/bin/bash
# round 2, run t: ncu --set full captures of the dominant launch of c3 / c4 / c5 (full per-GPU sizes) for traffic.json
mkdir -p gpurun_out
N="--no-e2e --no-cpu-baseline --no-configs --no-strong --no-eager --steps 1 --warmup 3"
for cfg in c3 c4 c5; do
  python bench.py $N --config $cfg > gpurun_out/r02t_plain_$cfg.json 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:kvc_fused -s 3 -c 2 -o gpurun_out/r02t_prof_$cfg -f python bench.py $N --config $cfg > gpurun_out/r02t_ncu_$cfg.log 2>&1
  tail -1 gpurun_out/r02t_ncu_$cfg.log
done
