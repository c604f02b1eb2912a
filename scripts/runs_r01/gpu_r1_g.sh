# full measurement table with the v2 kernel + ncu evidence for the long-context kernels
set -x
B="python bench.py --steps 10 --warmup 3 --no-e2e"
for c in c1 c2 c2_steady c3 c4 c5; do
  timeout 900 $B --config $c > gpurun_out/table_$c.json 2> gpurun_out/table_$c.err || tail -3 gpurun_out/table_$c.err
done
# B=1 latency regime (the reference's published regime): one decode stream
for c in c2 c2_steady c4; do
  timeout 600 $B --config $c --batch 1 --steps 20 --no-cpu-baseline > gpurun_out/table_${c}_b1.json 2> gpurun_out/table_${c}_b1.err || tail -3 gpurun_out/table_${c}_b1.err
done
N="--steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py $N --config c4 > gpurun_out/plain_c4.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:kvc_fused -s 3 -c 1 -o gpurun_out/prof_c4_v2 -f python bench.py $N --config c4 > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log
python bench.py $N --config c3 --batch 8 > gpurun_out/plain_c3.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:kvc_fused -s 3 -c 1 -o gpurun_out/prof_c3_v2 -f python bench.py $N --config c3 --batch 8 > gpurun_out/ncu_c3.log 2>&1
tail -2 gpurun_out/ncu_c3.log
