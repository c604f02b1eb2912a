#!/bin/bash
# LAB build of the device library: same sources with -DKVC_LAB, which turns the KVC_* environment overrides
# (launch shapes, CTA order, vote stage isolation) back on.  Output: csrc/libkvc_sm100a_lab.so — never loaded by the
# package; A/B scripts opt in with `KVC_LAB_LIBRARY=1 python scripts/...` (see scripts/lab_util.py).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
CSRC="$ROOT/cs3602-llm-inference-acceleration_b200/csrc"
mkdir -p "$CSRC/build_lab"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I $ROOT/include -DKVC_LAB"
pids=()
for part in 1 2 4 8 16 32; do
  nvcc $FLAGS -DKVC_PART=$part -c -o "$CSRC/build_lab/kvc_part$part.o" "$CSRC/kvc_sm100a.cu" 2> "$CSRC/build_lab/part$part.log" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$CSRC/libkvc_sm100a_lab.so" "$CSRC"/build_lab/kvc_part{1,2,4,8,16,32}.o -ldl
echo "built $CSRC/libkvc_sm100a_lab.so"
