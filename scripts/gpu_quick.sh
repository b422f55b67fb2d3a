mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/bench_layers_r2.py 8 > gpurun_out/r2_conv_layers_b8.txt 2>&1; grep -E "split|sum" gpurun_out/r2_conv_layers_b8.txt | awk '{print $1,$2,$3,$4}' | paste -sd' ' | fold -w 200
timeout 600 python bench.py --workload train --steps 20 --warmup 3 > gpurun_out/r2_bench_train.json 2> gpurun_out/r2_bench_train.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_train.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['kernel_ms_per_step'])"
