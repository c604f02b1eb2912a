set -x
timeout 300 python -m pytest tests/test_gpu_vote.py -x -q 2>&1 | tail -6
for v in 0 1; do for c in c2_vote; do
  KVC_VOTE_TMA=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --config $c > gpurun_out/vote_c2_$v.json 2>/dev/null; python -c "
import json,sys; d=json.loads(open('gpurun_out/vote_c2_$v.json').read().strip().splitlines()[-1]); print('VOTE tma=$v $c', d['us_per_step'], d['value'], d['roofline']['frac'], d['tensor_tflops'])"
done; done
