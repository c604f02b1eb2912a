set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
KVC_TMA_UPC=3 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 600 env "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err || tail -5 gpurun_out/$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$name.json").read().strip().splitlines()[-1])
    print("RESULT $name", d["value"], "GB/s", {k:(v["us_mean"], v["gbs"], v["frac_of_peak"]) for k,v in d["per_call"].items()})
except Exception as e: print("RESULT $name FAILED", e)
PY
}
for U in 1 2 4 8; do run c2_upc$U KVC_TMA_UPC=$U $B --config c2; done
for U in 1 4; do run c2_S8192_upc$U KVC_TMA_UPC=$U $B --config c2 --seq-len 8192 --batch 16; done
for U in 2 4; do run c2s_upc$U KVC_TMA_UPC=$U $B --config c2_steady; done
for U in 2 4; do run c3_upc$U KVC_TMA_UPC=$U $B --config c3; done
