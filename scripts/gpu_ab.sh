# Scratch A/B driver for gpurun: edit the commands below, then
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_ab.sh'
# Knobs the library reads per call: KVC_VOTE_SPLIT, KVC_VOTE_TS, KVC_VOTE_PEND, KVC_VOTE_TMA, KVC_FORCE_LDG, KVC_TMA_*.
run() { # config env...
  c=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --config $c > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); print('AB $c $*', d['us_per_step'], d['value'], d['roofline']['frac'])"
}
run c2
run c4_vote
run c4_vote KVC_VOTE_SPLIT=1 KVC_VOTE_TS=12 KVC_VOTE_PEND=3
