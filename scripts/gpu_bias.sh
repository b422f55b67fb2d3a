mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "wgrad" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_networks_gpu.py tests/test_entry_points_gpu.py -x -q -m gpu 2>&1 | tail -4
for bias in 0 1; do
MG_BIAS_IN_WGRAD=$bias MG_BENCH_NO_TORCH=1 timeout 600 python bench.py --workload train --steps 20 --warmup 3 --cpu-seconds 1 > gpurun_out/bias_$bias.json 2> gpurun_out/bias_$bias.err; python -c "
import json; d=json.load(open('gpurun_out/bias_$bias.json')); print('bias_in_wgrad', $bias, d['value'], d['ms_per_step'], d['e2e']['value'])"
done
