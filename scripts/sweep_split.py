"""Development aid: the split-operand kernel on the 64 x 64 / 32 x 32 layers with one or two CTAs per SM
(MG_SPLIT_2CTA_TILES = minimum tile count for the two-CTA shape; 0 = never), in one process."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200.networks import ops


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


for name, ci, co, H, B in [("D2.c2", 64, 64, 64, 8), ("D2.c2", 64, 64, 64, 16), ("D3.c1", 64, 80, 64, 8), ("D3.c1", 64, 80, 64, 16),
                           ("D3.c2", 80, 80, 32, 8), ("D3.c2", 80, 80, 32, 16), ("D4.c1", 80, 96, 32, 16), ("D4.c2", 96, 96, 16, 16)]:
    x = th.randn(B, ci, H, H, device="cuda").contiguous(memory_format=th.channels_last)
    dy = th.randn(B, co, H, H, device="cuda").contiguous(memory_format=th.channels_last)
    w = th.randn(co, ci, 3, 3, device="cuda")
    bias = th.randn(co, device="cuda")
    for op, fn in (("fprop3", lambda: ops.conv3x3(x, w, bias, lrelu=True, exact_w=True)), ("fprop2", lambda: ops.conv3x3(x, w)),
                   ("dgrad ", lambda: ops.conv3x3(dy, w, None, dgrad=True))):
        row = []
        for two in (0, 1, 148, 296):
            os.environ["MG_SPLIT_2CTA_TILES"] = str(two)
            row.append(f"{two}: {timed(fn):5.1f}")
        print(f"{name} B{B} {op}  2-CTA threshold " + "  ".join(row), flush=True)
