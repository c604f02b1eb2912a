#!/bin/bash
# round 2, run v: where the in-place c4 call (32K stored norms -> pool -> select -> slide 512 rows) spends its time
mkdir -p gpurun_out
python bench.py --mode slab --config c4 --steps 10 > gpurun_out/r02v_slab_c4.json 2> gpurun_out/r02v_slab_c4.err; echo "slab c4 rc=$?"; head -c 600 gpurun_out/r02v_slab_c4.json; echo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kvc_slab_compress -s 2 -c 1 -o gpurun_out/r02v_prof_slab_c4 -f python bench.py --mode slab --config c4 --batch 4 --steps 2 > gpurun_out/r02v_ncu.log 2>&1; tail -2 gpurun_out/r02v_ncu.log
