#!/bin/bash
# round 2, run k: full GPU suite (with the soak), padded-pitch per-channel capture, both bench arms, in-place slab lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest.log; tail -4 gpurun_out/r02k_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__cycles_active.avg.pct_of_peak_sustained_elapsed,dram__cycles_active.min.pct_of_peak_sustained_elapsed,dram__cycles_active.max.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k regex:kvc_fused -s 3 -c 1 --csv --log-file gpurun_out/r02k_chan_product_pad8.csv python scripts/stream_copy_control.py --only product --steps 2 --pad-rows 8 > /dev/null 2>&1
grep -v "^==" gpurun_out/r02k_chan_product_pad8.csv | cut -d, -f13-15 | tail -6
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02k_bench_ref.json 2> gpurun_out/r02k_bench_ref.err; echo "ref rc=$?"
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r02k_bench.err
for cfg in c2 c2_steady c3 c4 c5; do
  python bench.py --mode slab --config $cfg --steps 10 > gpurun_out/r02k_slab_$cfg.json 2> gpurun_out/r02k_slab_$cfg.err; echo "slab $cfg rc=$?"
done
python bench.py --mode slab --config c2_steady --batch 1 --steps 20 > gpurun_out/r02k_slab_c2_steady_b1.json 2>/dev/null
