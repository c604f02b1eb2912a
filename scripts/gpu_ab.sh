run() { # config batch env...
  c=$1; b=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --config $c --batch $b > gpurun_out/ab_$c.json 2> gpurun_out/ab.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$c.json').read().strip().splitlines()[-1]); print('AB $c B=$b $*', d['us_per_step'], d['roofline']['launch_us_min'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
}
OLD=KVC_LIBRARY=$PWD/scripts/ab/libkvc_old.so
run c4_vote 16 $OLD
run c4_vote 16
run c4_vote 16 $OLD
run c4_vote 16
run c2_vote 32 $OLD
run c2_vote 32
