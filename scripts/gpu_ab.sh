# Scratch A/B driver for gpurun: edit the commands below, then
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_ab.sh'
# Knobs the library reads per call: KVC_VOTE_SPLIT, KVC_VOTE_TS, KVC_VOTE_PEND, KVC_VOTE_TMA, KVC_VOTE_DEBUG (profiling
# only), KVC_FORCE_LDG, KVC_TMA_NT / _CTAS / _NSW / _UPC; KVC_LIBRARY=<path> loads another build of the library.
run() { # config batch env...
  c=$1; b=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --config $c --batch $b > gpurun_out/ab_$c.json 2> gpurun_out/ab.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$c.json').read().strip().splitlines()[-1]); print('AB $c B=$b $*', d['us_per_step'], d['value'], [(k, v['us_mean'], v['frac_of_peak']) for k,v in d['per_call'].items()])"
}
run c2 32
run c4_vote 16
run c2_vote 32
