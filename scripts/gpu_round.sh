#!/bin/bash
# One GPU-box session: full GPU test suite, smoke, both bench arms, in-graph kernel tables, per-layer tables, then (after
# those exited 0 without ncu) the ncu launch list of one eager iteration and full captures of the dominant kernels.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null
python bench.py --workload generate > gpurun_out/bench_generate.json 2>/dev/null
timeout 200 python scripts/profile_graph_step.py 2>&1 | grep -v Warn | grep "==\| ms " > gpurun_out/graph_step_kernels.txt
timeout 200 python scripts/bench_conv_layers.py 8 > gpurun_out/conv_layers_b8.txt 2>&1
timeout 200 python scripts/bench_pointwise.py > gpurun_out/pointwise.txt 2>&1
python scripts/one_train_iter.py 2 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_eager.csv python scripts/one_train_iter.py 2 > gpurun_out/ncu_iter.log 2>&1
python scripts/one_conv.py fprop 16 32 512 8 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3 -s 3 -c 1 -o gpurun_out/prof_conv_fprop_v5 -f python scripts/one_conv.py fprop 16 32 512 8 > gpurun_out/ncu_conv5.log 2>&1
python scripts/one_conv.py wgrad 32 32 256 8 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3_wgrad -s 3 -c 1 -o gpurun_out/prof_conv_wgrad_v6 -f python scripts/one_conv.py wgrad 32 32 256 8 > gpurun_out/ncu_wgrad6.log 2>&1
ls -la gpurun_out | tail -14
