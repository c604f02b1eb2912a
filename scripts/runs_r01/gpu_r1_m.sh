set -x
timeout 600 python -m pytest tests/test_gpu_vote.py -x -q 2>&1 | tail -4
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 900 $B --config c4_vote > gpurun_out/vote_c4.json 2> gpurun_out/vote_c4.err || tail -5 gpurun_out/vote_c4.err
python -c "
import json; d=json.loads(open('gpurun_out/vote_c4.json').read().strip().splitlines()[-1]); print('VOTE', d['us_per_step'], d['value'], d['tensor_tflops'])"
N="--steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py $N --config c4_vote --batch 4 > gpurun_out/plain_vote.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:kvc_snapkv_vote -s 3 -c 1 -o gpurun_out/prof_vote_c4_b4 -f python bench.py $N --config c4_vote --batch 4 > gpurun_out/ncu_vote.log 2>&1
tail -2 gpurun_out/ncu_vote.log
