set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python scripts/diag_copy.py
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 600 env "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err || tail -5 gpurun_out/$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$name.json").read().strip().splitlines()[-1])
    print("RESULT $name", d["value"], "GB/s", {k:(v["us_mean"], v["gbs"], v["frac_of_peak"]) for k,v in d["per_call"].items()})
except Exception as e: print("RESULT $name FAILED", e)
PY
}
run c4_tma          KVC_X=1 $B --config c4
run c4_tma_256      KVC_TMA_NT=256 $B --config c4
run c4_ldg          KVC_FORCE_LDG=1 $B --config c4
run c5_tma          KVC_X=1 $B --config c5
run c3_tma          KVC_X=1 $B --config c3
run c2_tma          KVC_X=1 $B --config c2
run c2_nsw4         KVC_TMA_NSW=4 $B --config c2
run c2_nsw6         KVC_TMA_NSW=6 $B --config c2
