timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { # config batch env...
  c=$1; b=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --config $c --batch $b > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); print('AB $c B=$b $*', d['us_per_step'], d['value'], [(k, v['us_mean'], v['frac_of_peak']) for k,v in d['per_call'].items()])"
}
run c2 32
run c2_steady 32
run c3 32
run c4_vote 16
run c2_vote 32
