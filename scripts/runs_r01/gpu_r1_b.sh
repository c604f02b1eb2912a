# v2 (TMA) kernel: parity, then per-config timing against the LDG form and staging variants
set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; echo "== $name"; timeout 600 env "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err || tail -5 gpurun_out/$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$name.json").read().strip().splitlines()[-1])
    print("$name", d["value"], "GB/s", {k:(v["us_mean"], v["gbs"], v["frac_of_peak"]) for k,v in d["per_call"].items()})
except Exception as e: print("$name FAILED", e)
PY
}
run c2_tma          KVC_X=1 $B --config c2
run c2_ldg          KVC_FORCE_LDG=1 $B --config c2
run c2_tma_c2       KVC_TMA_CTAS=2 $B --config c2
run c2_tma_512      KVC_TMA_NT=512 $B --config c2
run c2s_tma         KVC_X=1 $B --config c2_steady
run c2s_ldg         KVC_FORCE_LDG=1 $B --config c2_steady
run c3_tma          KVC_X=1 $B --config c3
run c3_ldg          KVC_FORCE_LDG=1 $B --config c3
run c4_tma          KVC_X=1 $B --config c4
run c4_tma_512      KVC_TMA_NT=512 $B --config c4
run c4_ldg          KVC_FORCE_LDG=1 $B --config c4
run c5_tma          KVC_X=1 $B --config c5
run c5_tma_512      KVC_TMA_NT=512 $B --config c5
run c5_ldg          KVC_FORCE_LDG=1 $B --config c5
run c1_tma          KVC_X=1 $B --config c1
run c1_ldg          KVC_FORCE_LDG=1 $B --config c1
