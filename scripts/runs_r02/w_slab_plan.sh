#!/bin/bash
# round 2, run w: in-place plan (kept indices alias the histogram, slide staging from the dead key array: 3 CTAs/SM at
# 32K rows) — slab tests incl. the soak, then the in-place bench lines
mkdir -p gpurun_out
python -m pytest tests/test_gpu_slab.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02w_pytest.log; tail -3 gpurun_out/r02w_pytest.log
for cfg in c2 c2_steady c3 c4 c5; do
  python bench.py --mode slab --config $cfg --steps 10 > gpurun_out/r02w_slab_$cfg.json 2> gpurun_out/r02w_slab_$cfg.err; echo "slab $cfg rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02w_slab_$cfg.json") if l.startswith("{")][-1])
print("$cfg", d["ms_per_step"], {k:(v["compress_us_mean"], v.get("append_us_mean")) for k,v in d["per_call"].items()})
PY
done
python bench.py --mode slab --config c2_steady --batch 1 --steps 20 > gpurun_out/r02w_slab_c2_steady_b1.json 2>/dev/null; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02w_slab_c2_steady_b1.json") if l.startswith("{")][-1])
print("c2_steady_b1", d["ms_per_step"], {k:(v["compress_us_mean"]) for k,v in d["per_call"].items()}, d.get("graph_replay"))
PY
