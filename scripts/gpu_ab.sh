F="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 python bench.py $F --config c4 > gpurun_out/r_c4.json 2>/dev/null
timeout 300 python bench.py $F --config c4_vote > gpurun_out/r_c4_vote.json 2>/dev/null
timeout 300 python bench.py $F --config c2_vote > gpurun_out/r_c2_vote.json 2>/dev/null
timeout 300 python bench.py --mode slab --config c4 --steps 10 > gpurun_out/r_slab_c4.json 2>/dev/null
for f in r_c4 r_c4_vote r_c2_vote; do python -c "
import json; d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('R $f', d['us_per_step'], d['value'], d['roofline']['frac'], d.get('tensor_tflops'))"; done
python -c "
import json; d=json.loads(open('gpurun_out/r_slab_c4.json').read().strip().splitlines()[-1]); print('R slab_c4', d['per_call'])"
