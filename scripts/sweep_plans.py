"""Development aid: sweep the launch-shape overrides of the convolution kernels on the layers where they matter.
usage: sweep_plans.py"""
import os, sys, subprocess
CHILD = r'''
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath("%s"))))
import torch as th
from musicgan_b200.networks import ops
ci, co, H, B, ups, pn = %d, %d, %d, %d, %d, %d
hin = H // 2 if ups else H
dt = th.float32 if ops.is_precise(H) else th.bfloat16
x = th.randn(B, ci, hin, hin, device="cuda").to(dt).contiguous(memory_format=th.channels_last)
w = th.randn(co, ci, 3, 3, device="cuda").requires_grad_(True)
bias = th.randn(co, device="cuda")
fn = lambda: ops.conv3x3(x, w, bias, lrelu=True, upsample_in=bool(ups), pixelnorm=bool(pn), split_w=True, exact_w=not pn)
try:
    fn(); th.cuda.synchronize()
    ws = {}
    g = th.cuda.CUDAGraph()
    with ops.capture_workspaces(ws), th.cuda.graph(g):
        for _ in range(20):
            fn()
    g.replay(); th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); th.cuda.synchronize()
    print("%%.1f" %% (a.elapsed_time(b) / 20 * 1e3))
except Exception as e:
    print("fail")
'''
def run(env, *shape):
    r = subprocess.run([sys.executable, "-c", CHILD % ((__file__,) + shape)], env=dict(os.environ, **env), capture_output=True, text=True)
    return r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "err"

BF16 = [("D0.c1", 16, 32, 512, 8, 0, 0), ("D0.c2", 32, 32, 256, 8, 0, 0), ("D1.c1", 32, 48, 256, 8, 0, 0), ("D1.c2", 48, 48, 128, 8, 0, 0),
        ("D2.c1", 48, 64, 128, 8, 0, 0), ("G5.c2", 64, 48, 128, 8, 1, 1), ("G6.c1", 48, 48, 128, 8, 0, 1), ("G6.c2", 48, 32, 256, 8, 1, 1),
        ("G7.c1", 32, 32, 256, 8, 0, 1), ("G7.c2", 32, 16, 512, 8, 1, 1)]
for name, *shape in BF16:
    row = [f"{name:6s} default {run({}, *shape):>6s}"]
    for occ in (1, 2, 3):
        for mb in (1, 2, 4):
            row.append(f"o{occ}m{mb} {run({'MG_CONV_OCC': str(occ), 'MG_CONV_MB': str(mb)}, *shape):>6s}")
    print("  ".join(row), flush=True)
SPLIT = [("D2.c2", 64, 64, 64, 16, 0, 0), ("D3.c1", 64, 80, 64, 16, 0, 0), ("D3.c2", 80, 80, 32, 16, 0, 0), ("D4.c1", 80, 96, 32, 16, 0, 0),
         ("D2.c2b8", 64, 64, 64, 8, 0, 0), ("D3.c1b8", 64, 80, 64, 8, 0, 0), ("G4.c2", 80, 64, 64, 8, 1, 1), ("G5.c1", 64, 64, 64, 8, 0, 1)]
for name, *shape in SPLIT:
    print(f"{name:8s} 1 CTA/SM {run({'MG_SPLIT_2CTA_TILES': '0'}, *shape):>6s}   2 CTAs/SM {run({'MG_SPLIT_2CTA_TILES': '1'}, *shape):>6s}", flush=True)
