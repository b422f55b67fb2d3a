#!/bin/bash
# One GPU-box session: full GPU test suite, both bench workloads, ncu launch lists (development aid).
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -5
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cat gpurun_out/bench_reference.json
