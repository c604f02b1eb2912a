timeout 600 python -m pytest tests/test_gpu_vote.py -x -q 2>&1 | tail -4
run() { # config batch env...
  c=$1; b=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --config $c --batch $b > gpurun_out/ab_$c.json 2> gpurun_out/ab.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$c.json').read().strip().splitlines()[-1]); print('AB $c B=$b $*', d['us_per_step'], d['value'], d['roofline']['frac'], d.get('tensor_tflops'))"
}
run c4_vote 16
run c2_vote 32
