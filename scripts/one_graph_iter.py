"""ncu target: ONE CUDA-graph critic step and ONE generator step of the train workload (stage 7, batch 8) between
cudaProfilerStart / Stop.  Use `ncu --profile-from-start off ...` so warm-up and capture are not profiled."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200 import bench_train
from musicgan_b200.graphed import GraphedSteps

dev = th.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
gen, disc = bench_train._build(7, 0, dev)
opt_g = th.optim.Adam(gen.parameters(), lr=1e-3, betas=(0.0, 0.9), capturable=True, fused=True)
opt_d = th.optim.Adam(disc.parameters(), lr=1e-3, betas=(0.0, 0.9), capturable=True, fused=True)
gs = GraphedSteps(gen, disc, opt_g, opt_d, B, 32, 512, 0.5)
x_real = th.rand(B, 2, 512, 512, device=dev) * 2 - 1
for _ in range(2):
    gs.critic_step(x_real); gs.generator_step()
th.cuda.synchronize()
th.cuda.cudart().cudaProfilerStart()
gs.critic_step(x_real)
gs.generator_step()
th.cuda.synchronize()
th.cuda.cudart().cudaProfilerStop()
print("ok")
