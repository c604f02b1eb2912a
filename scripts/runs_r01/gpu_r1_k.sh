set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for c in c2_steady c2 c3 c4 c5; do
  timeout 900 python bench.py --mode slab --config $c --steps 10 > gpurun_out/slab_$c.json 2> gpurun_out/slab_$c.err || tail -5 gpurun_out/slab_$c.err
done
timeout 600 python bench.py --mode slab --config c2_steady --batch 1 --steps 50 > gpurun_out/slab_c2_steady_b1.json 2> gpurun_out/slab_c2_steady_b1.err || tail -5 gpurun_out/slab_c2_steady_b1.err
