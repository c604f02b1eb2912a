set -x
timeout 600 python -m pytest tests/test_gpu_vote.py tests/test_gpu_slab.py -x -q 2>&1 | tail -4
for rep in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --config c4_vote 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VOTE c4 rep$rep', d['us_per_step'], d['value'], d['roofline']['frac'], d['tensor_tflops'])"
done
timeout 300 python bench.py --mode slab --config c4 --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SLAB c4', d['per_call'])"
