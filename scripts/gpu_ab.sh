timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { # config batch env...
  c=$1; b=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --config $c --batch $b > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); print('AB $c B=$b $*', d['us_per_step'], d['value'], d['roofline']['frac'])"
}
run c4 16
run c4 1
run c4_vote 16
timeout 300 python bench.py --mode slab --config c4 --steps 10 > gpurun_out/slab_c4_new.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/slab_c4_new.json').read().strip().splitlines()[-1]); print('SLAB c4', d['per_call'])"
