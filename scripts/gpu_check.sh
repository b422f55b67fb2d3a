mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_umma_probe_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m pytest tests/test_networks_gpu.py -x -q -m gpu 2>&1 | tail -3
MG_BENCH_NO_TORCH=1 timeout 600 python bench.py --workload train --steps 20 --warmup 3 --cpu-seconds 1 > gpurun_out/check_train.json 2> gpurun_out/check_train.err; python -c "
import json; d=json.load(open('gpurun_out/check_train.json')); print('train', d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['kernel_ms_per_step']); print({k: d['roofline'][k] for k in ('frac','per_launch_bound_frac','step_frac')})"
timeout 300 python scripts/bench_layers_r2.py 8 > gpurun_out/check_layers.txt 2>&1; tail -5 gpurun_out/check_layers.txt
