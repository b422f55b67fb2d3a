"""Development aid: time one weight-gradient launch with parts of k_conv3x3_wgrad switched off (MG_WGRAD_ABLATE bits:
1 no copies, 2 no MMAs, 4 no operand staging (ldmatrix / tcgen05.st)).  usage: ablate_wgrad.py Cin Cout H B"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200.networks import ops
ci, co, H, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
x = th.randn(B, ci, H, H, device="cuda").bfloat16().contiguous(memory_format=th.channels_last)
dy = th.randn(B, co, H, H, device="cuda").bfloat16().contiguous(memory_format=th.channels_last)
out = th.empty(co, ci, 3, 3, device="cuda")
fn = lambda: ops.conv3x3_wgrad(dy, x, out=out)
for abl, bufs in ((0, 4), (7, 4), (15, 4), (1, 4), (9, 4)):
    os.environ["MG_WGRAD_ABLATE"] = str(abl); os.environ["MG_WGRAD_ABUFS"] = str(bufs)
    fn(); th.cuda.synchronize()
    ws = {}
    g = th.cuda.CUDAGraph()
    with ops.capture_workspaces(ws), th.cuda.graph(g):
        for _ in range(10):
            fn()
    g.replay(); th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); th.cuda.synchronize()
    print(f"wgrad {' '.join(sys.argv[1:5]):18s} ablate {abl} a_bufs<={bufs}: {a.elapsed_time(b) / 10 * 1e3:7.1f} us (kernel + reduce)", flush=True)
