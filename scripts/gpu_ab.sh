run() { # config batch env...
  c=$1; b=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --config $c --batch $b > gpurun_out/ab_$c.json 2> gpurun_out/ab.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$c.json').read().strip().splitlines()[-1]); print('AB $c B=$b $*', d['us_per_step'], d['value'], [(k, v['us_mean'], v['frac_of_peak']) for k,v in d['per_call'].items()])"
}
EF=KVC_LIBRARY=$PWD/scripts/ab/libkvc_ef.so
run c2 32
run c2 32 $EF
run c3 32
run c3 32 $EF
run c4 16
run c4 16 $EF
run c5 8
run c5 8 $EF
