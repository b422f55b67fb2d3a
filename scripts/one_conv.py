"""Run a few launches of one convolution shape (target for ncu captures).  usage: one_conv.py fprop|dgrad|wgrad Cin Cout H B [ups]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200.networks import ops
op, ci, co, H, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
ups = len(sys.argv) > 6 and sys.argv[6] == "ups"
hin = H // 2 if ups else H
x = th.randn(B, ci, hin, hin, device="cuda").bfloat16().contiguous(memory_format=th.channels_last)
dy = th.randn(B, co, H, H, device="cuda").bfloat16().contiguous(memory_format=th.channels_last)
w = th.randn(co, ci, 3, 3, device="cuda")
bias = th.randn(co, device="cuda")
for _ in range(6):
    if op == "fprop":
        ops.conv3x3(x, w, bias, lrelu=True, upsample_in=ups)
    elif op == "dgrad":
        ops.conv3x3(dy, w, None, dgrad=True)
    else:
        ops.conv3x3_wgrad(dy, x, upsample_in=ups)
th.cuda.synchronize()
print("ok")
