#!/bin/bash
# Round-2 GPU session: full GPU suite (complete log), smoke, the bench arms and workloads, in-graph kernel table.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -s > gpurun_out/r2_tests_full.log 2>&1; tail -15 gpurun_out/r2_tests_full.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 400 gpurun_out/r2_bench_default.err; head -c 600 gpurun_out/r2_bench_default.json
timeout 600 python bench.py --impl torch_cuda --steps 5 --warmup 3 > gpurun_out/r2_bench_torch.json 2> gpurun_out/r2_bench_torch.err; head -c 900 gpurun_out/r2_bench_torch.json
timeout 900 python bench.py --workload sweep --batch 64 --steps 5 --warmup 3 > gpurun_out/r2_bench_sweep.json 2> gpurun_out/r2_bench_sweep.err; tail -c 300 gpurun_out/r2_bench_sweep.err; head -c 300 gpurun_out/r2_bench_sweep.json
timeout 600 python bench.py --workload generate --steps 5 > gpurun_out/r2_bench_generate.json 2> gpurun_out/r2_bench_generate.err; tail -c 300 gpurun_out/r2_bench_generate.err; head -c 300 gpurun_out/r2_bench_generate.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; head -c 400 gpurun_out/r2_bench_reference.json
timeout 300 python scripts/profile_graph_step.py 2>&1 | grep -v Warn | grep "==\| ms " > gpurun_out/r2_graph_step_kernels.txt; tail -30 gpurun_out/r2_graph_step_kernels.txt
