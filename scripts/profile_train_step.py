"""Kernel-time breakdown of the train workload with torch.profiler (development aid, GPU box only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from torch.profiler import profile, ProfilerActivity
from musicgan_b200 import bench_train

dev = th.device("cuda", 0)
gen, disc = bench_train._build(7, 0, dev)
opt_g = th.optim.Adam(gen.parameters(), lr=1e-3, betas=(0.0, 0.9))
opt_d = th.optim.Adam(disc.parameters(), lr=1e-3, betas=(0.0, 0.9))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
x_real = (th.rand(B, 2, 512, 512, device=dev) * 2 - 1)


def d_step():
    z = th.randn(B, 32, 2, 2, device=dev)
    with th.no_grad():
        x_fake = gen(z, 0.5)
    ob = disc(th.cat([x_real, x_fake], 0), 0.5)
    d_loss = -(ob[:B].mean() - ob[B:].mean())
    gp = disc.gradient_penalty(x_real, x_fake, 0.5)
    gen.zero_grad(); disc.zero_grad()
    (d_loss + gp).backward()
    opt_d.step()


def g_step():
    z = th.randn(B, 32, 2, 2, device=dev)
    g_loss = -disc(gen(z, 0.5), 0.5).mean()
    gen.zero_grad(); disc.zero_grad()
    g_loss.backward()
    opt_g.step()


for _ in range(2):
    d_step(); g_step()
th.cuda.synchronize()
for name, fn in (("critic step", d_step), ("generator step", g_step)):
    import time
    th.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        fn()
    th.cuda.synchronize(); print(f"{name}: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms wall")
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        fn()
        th.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
