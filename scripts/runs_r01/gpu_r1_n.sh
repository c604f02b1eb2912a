set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "full_baseline_size or slab" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_slab.py -x -q 2>&1 | tail -3
for sl in 16 32; do
  timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-slab $sl > gpurun_out/e2e_slab$sl.json 2> gpurun_out/e2e_slab$sl.err
  python -c "
import json; d=json.loads(open('gpurun_out/e2e_slab$sl.json').read().strip().splitlines()[-1]); e=d['e2e']; print('E2E slab $sl', e['value'], e['ms_per_step'], e['alternative']['value'])"
done
timeout 600 python bench.py --mode slab --config c2_steady --batch 1 --steps 50 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SLAB B1', d['wall_ms_per_step'], d['per_call'])"
