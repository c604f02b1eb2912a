N="--steps 1 --warmup 3 --no-e2e --no-cpu-baseline --config c2_vote --batch 8"
python bench.py $N > gpurun_out/plain_vote.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kvc_snapkv_vote -s 3 -c 1 -o gpurun_out/prof_vote_c2 -f python bench.py $N > gpurun_out/ncu_vote.log 2>&1
tail -2 gpurun_out/ncu_vote.log
