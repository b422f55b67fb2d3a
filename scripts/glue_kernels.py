"""Development aid: full names and durations of the non-library kernels inside the critic graph replay."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from torch.profiler import profile, ProfilerActivity
from musicgan_b200 import bench_train
from musicgan_b200.graphed import GraphedSteps
dev = th.device("cuda", 0)
gen, disc = bench_train._build(7, 0, dev)
opt_g = th.optim.Adam(gen.parameters(), lr=1e-3, betas=(0.0, 0.9), capturable=True, fused=True)
opt_d = th.optim.Adam(disc.parameters(), lr=1e-3, betas=(0.0, 0.9), capturable=True, fused=True)
gs = GraphedSteps(gen, disc, opt_g, opt_d, 8, 32, 512, 0.5)
x_real = th.rand(8, 2, 512, 512, device=dev) * 2 - 1
for _ in range(3):
    gs.critic_step(x_real)
th.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gs.critic_step(x_real); th.cuda.synchronize()
evs = sorted((e for e in prof.events() if e.device_type == th.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
prev = None
for e in evs:
    d = e.time_range.end - e.time_range.start
    if "mg::" not in e.name and d >= 4:
        print(f"{d:7.1f} us  {e.name[:150]}   [after {prev[:40] if prev else None}]")
    prev = e.name.split('(')[0][-40:]
