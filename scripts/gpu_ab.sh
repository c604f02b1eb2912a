timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --config c4_vote 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RES c4_vote', d['us_per_step'], d['roofline']['frac'])"
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE ok')" 2>&1 | tail -2
