#!/bin/bash
# round 2, run u: NVTX ranges in the C entry points — suite, per-call host overhead at batch 1 (must not move), and the
# launches ncu attributes to each range (--nvtx --nvtx-include "<range>/")
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02u_pytest.log; tail -2 gpurun_out/r02u_pytest.log
python scripts/host_overhead.py > gpurun_out/r02u_host_overhead.log 2>&1; head -14 gpurun_out/r02u_host_overhead.log
for r in kvc_compress_layers_ws kvc_slab_append kvc_slab_compress kvc_snapkv_vote_compress; do
  timeout 600 ncu --nvtx --nvtx-include "$r/" --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02u_nvtx_$r.csv python scripts/sanitize.py > /dev/null 2>&1
  echo "$r: $(grep -c kvc_ gpurun_out/r02u_nvtx_$r.csv) launches; kernels: $(grep -o 'kvc_[a-z_]*kernel' gpurun_out/r02u_nvtx_$r.csv | sort | uniq -c | tr '\n' ' ')"
done
