"""Development aid: sweep the launch-shape overrides of the bf16 convolution kernel (MG_CONV_OCC CTAs per SM, MG_CONV_MB
blocks per pipeline step, MG_CONV_ACC accumulators) on the stage-7 layers, in one process.  usage: sweep_plans.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200.networks import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
LAYERS = [("D0.c1", 16, 32, 512, 0, 0), ("D0.c2", 32, 32, 256, 0, 0), ("D1.c1", 32, 48, 256, 0, 0), ("D1.c2", 48, 48, 128, 0, 0),
          ("D2.c1", 48, 64, 128, 0, 0), ("G5.c2", 64, 48, 128, 1, 1), ("G6.c1", 48, 48, 128, 0, 1), ("G6.c2", 48, 32, 256, 1, 1),
          ("G7.c1", 32, 32, 256, 0, 1), ("G7.c2", 32, 16, 512, 1, 1)]


def timed(fn):
    fn(); th.cuda.synchronize()
    ws = {}
    g = th.cuda.CUDAGraph()
    with ops.capture_workspaces(ws), th.cuda.graph(g):
        for _ in range(10):
            fn()
    g.replay(); th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); th.cuda.synchronize()
    return a.elapsed_time(b) / 10 * 1e3


for name, ci, co, H, ups, pn in LAYERS:
    hin = H // 2 if ups else H
    x = th.randn(B, ci, hin, hin, device="cuda").bfloat16().contiguous(memory_format=th.channels_last)
    dy = th.randn(B, co, H, H, device="cuda").bfloat16().contiguous(memory_format=th.channels_last)
    w = th.randn(co, ci, 3, 3, device="cuda")
    bias = th.randn(co, device="cuda")
    for op in ("fprop", "dgrad"):
        if op == "dgrad" and ups:
            fn = lambda: ops.conv3x3(dy, w, None, dgrad=True)
        elif op == "dgrad":
            fn = lambda: ops.conv3x3(dy, w, None, dgrad=True)
        else:
            fn = lambda: ops.conv3x3(x, w, bias, lrelu=True, upsample_in=bool(ups), pixelnorm=bool(pn), split_w=True)
        for k in ("MG_CONV_OCC", "MG_CONV_MB", "MG_CONV_ACC"):
            os.environ.pop(k, None)
        row = [f"{name} {op} default {timed(fn):6.1f}"]
        best = (1e9, "")
        for occ in (1, 2, 3, 4):
            for mb in (1, 2, 4):
                os.environ["MG_CONV_OCC"], os.environ["MG_CONV_MB"] = str(occ), str(mb)
                t = timed(fn)
                row.append(f"o{occ}m{mb} {t:5.1f}")
                best = min(best, (t, f"o{occ}m{mb}"))
        print("  ".join(row) + f"   best {best[1]} {best[0]:.1f}", flush=True)
