mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dp_gpu.py -q -m gpu -s 2>&1 | grep -v Warn | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; tail -c 600 gpurun_out/r2_bench_2gpu.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_2gpu.json')); print('2gpu train', d['value'], d['ms_per_step'], d['e2e']['value']); s=d['secondary']; print('2gpu transform', s['value'], s['ms_per_step'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload train --batch 64 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu_b64.json 2> gpurun_out/r2_bench_2gpu_b64.err; tail -c 300 gpurun_out/r2_bench_2gpu_b64.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_2gpu_b64.json')); print('2gpu b64', d['value'], d['ms_per_step'], d['e2e']['value'], d['config'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --workload transform --steps 2 --warmup 1 2>/dev/null | head -c 300
