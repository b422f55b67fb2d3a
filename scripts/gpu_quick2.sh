mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_networks_gpu.py -q -m gpu -s -k "inference_mode or forward_vs" 2>&1 | grep -E "rel-L2|passed|failed"
timeout 300 python -m pytest tests/test_entry_points_gpu.py -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --workload generate --steps 5 > gpurun_out/r2_bench_generate.json 2> gpurun_out/r2_bench_generate.err; python -c "
import json; g=json.load(open('gpurun_out/r2_bench_generate.json')); print(g['value'], g['ms_per_step'], g['e2e']['value'], g['roofline']['frac'], g['roofline']['kernel_ms_per_step'])"
timeout 600 python bench.py --workload generate --nb-vec 10 --gen-clips 64 --steps 3 > gpurun_out/r2_bench_generate_nv10.json 2> gpurun_out/r2_bench_generate_nv10.err; python -c "
import json; g=json.load(open('gpurun_out/r2_bench_generate_nv10.json')); print(g['value'], g['ms_per_step'], g['e2e']['value'], g['roofline']['frac'], g['config']['workload'])"
