set -x
KVC_VOTE_W8=1 timeout 600 python -m pytest tests/test_gpu_vote.py -x -q 2>&1 | tail -4
for rep in 1 2; do for v in 0 1; do for c in c4_vote c2_vote; do
  KVC_VOTE_W8=$v timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --config $c 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VOTE w8=$v $c rep$rep', d['us_per_step'], d['value'], d['roofline']['frac'])"
done; done; done
