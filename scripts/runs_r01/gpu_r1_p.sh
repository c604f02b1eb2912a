set -x
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
python bench.py --mode slab --config c5 --steps 1 > gpurun_out/plain_slab_c5.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:kvc_slab_compress -s 6 -c 2 -o gpurun_out/prof_slab_c5_after -f python bench.py --mode slab --config c5 --steps 1 > gpurun_out/ncu_slab_c5.log 2>&1
tail -2 gpurun_out/ncu_slab_c5.log
