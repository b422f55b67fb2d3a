mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_networks_gpu.py tests/test_entry_points_gpu.py -x -q -m gpu 2>&1 | tail -5
for lanes in 0 1; do
MG_WGRAD_LANES=$lanes MG_BENCH_NO_TORCH=1 timeout 600 python bench.py --workload train --steps 20 --warmup 3 --cpu-seconds 1 > gpurun_out/lane_$lanes.json 2> gpurun_out/lane_$lanes.err; python -c "
import json; d=json.load(open('gpurun_out/lane_$lanes.json')); print('lanes', $lanes, d['value'], d['ms_per_step'], d['e2e']['value'])"
done
