set -x
for rep in 1 2; do
for v in A B; do
  if [ $v = B ]; then export KVC_LIBRARY=$PWD/cs3602-llm-inference-acceleration_b200/csrc/libkvc_variantB.so; else unset KVC_LIBRARY; fi
  timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --config c4_vote 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VOTE $v rep$rep', d['us_per_step'], d['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['library'])"
done; done
