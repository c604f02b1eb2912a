set -x
for c in c2_steady; do
  timeout 900 python bench.py --mode slab --config $c --steps 10 > gpurun_out/slab_$c.json 2> gpurun_out/slab_$c.err || tail -5 gpurun_out/slab_$c.err
done
timeout 600 python bench.py --mode slab --config c2_steady --batch 1 --steps 50 > gpurun_out/slab_c2_steady_b1.json 2> gpurun_out/slab_c2_steady_b1.err || tail -5 gpurun_out/slab_c2_steady_b1.err
KVC_TMA_NT=256 timeout 600 python bench.py --mode slab --config c5 --steps 10 > gpurun_out/slab_c5_nt256.json 2> gpurun_out/slab_c5_nt256.err
KVC_TMA_NT=256 timeout 600 python bench.py --mode slab --config c4 --steps 10 > gpurun_out/slab_c4_nt256.json 2> gpurun_out/slab_c4_nt256.err
python bench.py --mode slab --config c5 --steps 1 > gpurun_out/plain_slab_c5.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:kvc_slab_compress -s 6 -c 2 -o gpurun_out/prof_slab_c5 -f python bench.py --mode slab --config c5 --steps 1 > gpurun_out/ncu_slab_c5.log 2>&1
tail -2 gpurun_out/ncu_slab_c5.log
