for k in 10 50; do
timeout 300 python bench.py --steps $k --warmup 5 --no-e2e --config c1 > gpurun_out/ab.json 2> gpurun_out/ab.err
python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); print('C1 steps=$k', d['us_per_step'], d['value'], d['roofline']['frac'], d['roofline']['launch_us_min'], d['clocks'], d['cpu_baseline']['value'])"
done
cp gpurun_out/ab.json gpurun_out/bench_c1_k50.json
