#!/bin/bash
# One GPU-box session: full GPU test suite, smoke, both bench arms, then (after those exited 0 without ncu) the ncu
# launch list of one graph iteration and full captures of the two dominant kernels (development aid).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cat gpurun_out/bench_reference.json
python scripts/one_graph_iter.py && timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_graph.csv python scripts/one_graph_iter.py > gpurun_out/ncu_graph.log 2>&1
python scripts/one_conv.py fprop 16 32 512 8 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3 -s 3 -c 1 -o gpurun_out/prof_conv_fprop_v4 -f python scripts/one_conv.py fprop 16 32 512 8 > gpurun_out/ncu_conv4.log 2>&1
python scripts/one_conv.py wgrad 32 32 256 8 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3_wgrad -s 3 -c 1 -o gpurun_out/prof_conv_wgrad_v4 -f python scripts/one_conv.py wgrad 32 32 256 8 > gpurun_out/ncu_wgrad4.log 2>&1
ls -la gpurun_out | tail -8
