#!/bin/bash
# round 2, run a: refactored library (generic widths, stored norms, pinned slab, fused vote) — tests + smoke
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -15 gpurun_out/r02a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02a_smoke.log
tail -5 gpurun_out/r02a_smoke.log
