set -x
timeout 300 python -m pytest tests/test_gpu_vote.py -x -q 2>&1 | tail -4
for rep in 1 2; do for v in A B; do
  if [ $v = A ]; then export KVC_LIBRARY=$PWD/cs3602-llm-inference-acceleration_b200/csrc/libkvc_variantA.so; else unset KVC_LIBRARY; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --config c4_vote 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VOTE $v rep$rep', d['us_per_step'], d['value'], d['roofline']['frac'])"
done; done
