#!/bin/bash
# End-of-round check on a GPU box: full GPU suite, smoke, both bench arms, generate workload, tables.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 200 gpurun_out/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null
python bench.py --workload generate > gpurun_out/bench_generate.json 2>/dev/null
timeout 200 python scripts/bench_conv_layers.py 8 > gpurun_out/conv_layers_b8.txt 2>&1
timeout 200 python scripts/bench_pointwise.py > gpurun_out/pointwise.txt 2>&1
MG_PDL=0 MG_TWO_STREAMS=0 timeout 200 python scripts/profile_graph_step.py 2>&1 | grep -v Warn | grep "==\| ms " > gpurun_out/graph_step_kernels_serial.txt
timeout 200 python scripts/profile_graph_step.py 2>&1 | grep -v Warn | grep "==" > gpurun_out/graph_step_spans.txt
wc -c gpurun_out/bench_default.json gpurun_out/bench_reference.json gpurun_out/bench_generate.json
