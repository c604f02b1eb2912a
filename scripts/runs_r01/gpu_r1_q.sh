set -x
N="--steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py $N --config c4_vote --batch 4 > gpurun_out/plain_vote.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:kvc_snapkv_vote -s 3 -c 1 -o gpurun_out/prof_vote_tma -f python bench.py $N --config c4_vote --batch 4 > gpurun_out/ncu_vote.log 2>&1
tail -2 gpurun_out/ncu_vote.log
