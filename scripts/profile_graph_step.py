"""Development aid (GPU box): kernel timeline of the CUDA-graph replays of the train workload (torch.profiler / CUPTI
kernel records): per-kernel GPU time inside the graph, GPU busy time and the idle gaps between kernels."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from torch.profiler import profile, ProfilerActivity
from musicgan_b200 import bench_train
from musicgan_b200.graphed import GraphedSteps

dev = th.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
gen, disc = bench_train._build(7, 0, dev)
opt_g = th.optim.Adam(gen.parameters(), lr=1e-3, betas=(0.0, 0.9), capturable=True, fused=True)
opt_d = th.optim.Adam(disc.parameters(), lr=1e-3, betas=(0.0, 0.9), capturable=True, fused=True)
gs = GraphedSteps(gen, disc, opt_g, opt_d, B, 32, 512, 0.5)
x_real = th.rand(B, 2, 512, 512, device=dev) * 2 - 1
for _ in range(3):
    gs.critic_step(x_real); gs.generator_step()
th.cuda.synchronize()
for name, fn in (("critic graph", lambda: gs.critic_step(x_real)), ("generator graph", gs.generator_step)):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        th.cuda.synchronize()
    evs = sorted((e for e in prof.events() if e.device_type == th.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    if not evs:
        print(name, "no kernel records"); continue
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    busy = sum(e.time_range.end - e.time_range.start for e in evs)
    agg = collections.defaultdict(lambda: [0.0, 0])
    for e in evs:
        k = e.name.split("(")[0][-60:]
        agg[k][0] += e.time_range.end - e.time_range.start; agg[k][1] += 1
    print(f"== {name}: {len(evs)} kernels, span {(t1 - t0) / 1e3:.3f} ms, sum of kernel times {busy / 1e3:.3f} ms, "
          f"gaps {(t1 - t0 - busy) / 1e3:.3f} ms")
    for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]:
        print(f"   {t / 1e3:8.3f} ms  {n:4d} x {t / n:7.1f} us  {k}")
    for pat in ("k_conv3x3<", "k_conv3x3_wgrad", "k_lrelu_bwd", "k_wgrad_reduce"):
        d = sorted(round(e.time_range.end - e.time_range.start, 1) for e in evs if pat in e.name)
        print(f"   durations of {pat} (us): {d}")
