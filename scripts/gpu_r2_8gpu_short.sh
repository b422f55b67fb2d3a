mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; tail -c 200 gpurun_out/$2.err | tr '\n' ' '; head -c 260 gpurun_out/$2.json; echo; }
run 29521 r2_bench_8gpu --steps 20 --warmup 3
run 29522 r2_bench_8gpu_b64 --workload train --batch 64 --steps 5 --warmup 3
