#!/bin/bash
# round 2, run c: compute-sanitizer over every kernel family, the streaming_llm copy control, ncu captures
# NOTE: the KVC_TMA_ORDER knob swept below was removed after this run (no effect measured: profiles/r02_stream_copy_control.json); compute-sanitizer turned out to be closed on this pool.
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
python scripts/sanitize.py > gpurun_out/r02c_sanitize_plain.log 2>&1; echo "plain rc=$?" | tee -a gpurun_out/r02c_sanitize_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize.py > gpurun_out/r02c_sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/r02c_sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize.py ok|Error|error" gpurun_out/r02c_sanitize_$tool.log | head -8
done
# streaming_llm at c2: product kernel vs plain-copy controls over the same address set
python scripts/stream_copy_control.py --out gpurun_out/r02c_stream_control.json > gpurun_out/r02c_stream_control.log 2>&1
tail -8 gpurun_out/r02c_stream_control.log
for order in 0 1 2; do
  KVC_LAB_LIBRARY=1 KVC_TMA_ORDER=$order python scripts/stream_copy_control.py --out gpurun_out/r02c_stream_control.json >> gpurun_out/r02c_stream_control.log 2>&1
  tail -1 gpurun_out/r02c_stream_control.log
done
# per-channel DRAM activity of the product kernel and the LDG control
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__cycles_active.avg.pct_of_peak_sustained_elapsed,dram__cycles_active.min.pct_of_peak_sustained_elapsed,dram__cycles_active.max.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k regex:kvc_fused -s 3 -c 1 --csv --log-file gpurun_out/r02c_chan_product.csv python scripts/stream_copy_control.py --only product --steps 2 > /dev/null 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:copyctl -s 6 -c 2 --csv --log-file gpurun_out/r02c_chan_ldg.csv python scripts/stream_copy_control.py --only ldg_4cta --steps 1 > /dev/null 2>&1
for order in 1 2; do
  KVC_LAB_LIBRARY=1 KVC_TMA_ORDER=$order timeout 600 ncu --metrics $M --clock-control none -k regex:kvc_fused -s 3 -c 1 --csv --log-file gpurun_out/r02c_chan_order$order.csv python scripts/stream_copy_control.py --steps 2 > /dev/null 2>&1
done
tail -2 gpurun_out/r02c_chan_*.csv | cut -c1-400
# launch list + full capture of the headline config (reduced command: the c2 timing leg only)
N="--no-e2e --no-cpu-baseline --no-configs --no-strong --no-eager --steps 2 --warmup 3"
python bench.py $N > gpurun_out/r02c_bench_c2_plain.json 2> gpurun_out/r02c_bench_c2_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kvc_ -c 400 --csv --log-file gpurun_out/r02c_launches_c2.csv python bench.py $N > gpurun_out/r02c_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kvc_fused -s 6 -c 2 -o gpurun_out/r02c_prof_c2 -f python bench.py $N > gpurun_out/r02c_ncu_full.log 2>&1
tail -2 gpurun_out/r02c_ncu_full.log
# the fused vote at the c4 shape (B = 4: one wave per SM takes ~3 ms under ncu replay)
V="--no-e2e --no-cpu-baseline --no-configs --no-strong --no-eager --steps 1 --warmup 3 --config c4_vote --batch 4"
python bench.py $V > gpurun_out/r02c_bench_vote_b4.json 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:kvc_snapkv_vote -s 3 -c 1 -o gpurun_out/r02c_prof_vote -f python bench.py $V > gpurun_out/r02c_ncu_vote.log 2>&1
tail -2 gpurun_out/r02c_ncu_vote.log
ls -la gpurun_out | tail -30
