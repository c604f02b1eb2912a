# what the driver runs at round end, in its order
set -x
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err ) 2>&1 | grep real
( time python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/final_c2.json 2> gpurun_out/final_c2.err ) 2>&1 | grep real
tail -c 400 gpurun_out/final_ref.json; echo; tail -3 gpurun_out/final_c2.err
