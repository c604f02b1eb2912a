set -x
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 600 env "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err || tail -5 gpurun_out/$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$name.json").read().strip().splitlines()[-1])
    print("RESULT $name", d["value"], "GB/s", {k:(v["us_mean"], v["gbs"], v["frac_of_peak"]) for k,v in d["per_call"].items()}, d["clocks"])
except Exception as e: print("RESULT $name FAILED", e)
PY
}
for S in 640 768 1024 2048 8192; do run c2_S$S KVC_X=1 $B --config c2 --seq-len $S --batch 16; done
run c2_b16 KVC_X=1 $B --config c2 --batch 16
# ncu: launch list + full capture of both c2 launches at the full config
$B --config c2 --steps 2 > gpurun_out/plain_c2.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_v2.csv -k regex:kvc_ $B --config c2 --steps 2 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
$B --config c2 --steps 1 > gpurun_out/plain_c2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:kvc_fused -s 6 -c 2 -o gpurun_out/prof_c2_v2 -f $B --config c2 --steps 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
