set -x
for v in 0 1 2; do
  if [ $v = 0 ]; then unset KVC_TMA_CTAS; else export KVC_TMA_CTAS=$v; fi
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('E2E ctas=$v', d['value'], e['value'], e['ms_per_step'], e['how'][:9], e['alternative']['value'])"
done
