"""One eager critic step + one generator step at stage 7, batch 8 (target for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200 import bench_train, train_step
dev = th.device("cuda", 0)
gen, disc = bench_train._build(7, 0, dev)
og = th.optim.Adam(gen.parameters(), lr=1e-3, betas=(0.0, 0.9))
od = th.optim.Adam(disc.parameters(), lr=1e-3, betas=(0.0, 0.9))
B = 8
x_real = th.rand(B, 2, 512, 512, device=dev) * 2 - 1
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for i in range(n):
    z = th.randn(B, 32, 2, 2, device=dev)
    train_step.critic_step(gen, disc, od, z, x_real, 0.5)
    train_step.generator_step(gen, disc, og, th.randn(B, 32, 2, 2, device=dev), 0.5)
th.cuda.synchronize()
print("ok")
