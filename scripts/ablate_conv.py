"""Development aid: time one bf16 convolution with parts of k_conv3x3 switched off (MG_CONV_ABLATE bits: 1 no halo copies,
2 no MMAs, 4 no epilogue, 8 no stores, 32 all taps read 128-byte aligned core matrices (wrong results: what would aligned
operand reads buy?)).  usage: ablate_conv.py fprop|dgrad Cin Cout H B"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200.networks import ops
op, ci, co, H, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
x = th.randn(B, ci, H, H, device="cuda").bfloat16().contiguous(memory_format=th.channels_last)
dy = th.randn(B, co, H, H, device="cuda").bfloat16().contiguous(memory_format=th.channels_last)
w = th.randn(co, ci, 3, 3, device="cuda")
bias = th.randn(co, device="cuda")
fn = (lambda: ops.conv3x3(x, w, bias, lrelu=True, split_w=True)) if op == "fprop" else (lambda: ops.conv3x3(dy, w, None, dgrad=True))
for abl in [int(a) for a in os.environ.get('ABL', '0,2,1,4,3,8,11,19,9,10').split(',')]:
    os.environ["MG_CONV_ABLATE"] = str(abl)
    fn(); th.cuda.synchronize()
    ws = {}
    g = th.cuda.CUDAGraph()
    with ops.capture_workspaces(ws), th.cuda.graph(g):
        for _ in range(20):
            fn()
    g.replay(); th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); th.cuda.synchronize()
    print(f"{' '.join(sys.argv[1:6]):26s} ablate {abl:2d}: {a.elapsed_time(b) / 20 * 1e3:7.1f} us", flush=True)
