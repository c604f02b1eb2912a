set -x
timeout 300 python -m pytest tests/test_gpu_vote.py -x -q 2>&1 | tail -8
for v in 0 1; do for c in c4_vote; do
  KVC_VOTE_TMA=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --config $c 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VOTE tma=$v $c', d['us_per_step'], d['value'], d['roofline']['frac'], d['tensor_tflops'])"
done; done
for dbg in 1 2; do
  KVC_VOTE_DEBUG=$dbg timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --config c4_vote 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VOTE tma dbg=$dbg', d['us_per_step'], d['value'])"
done
