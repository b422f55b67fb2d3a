"""Run a few launches of one convolution shape the way the networks run it in round 2 (target for ncu captures):
fp32 activations + split-operand kernel when the output height is <= ops.PRECISE_MAX_RES, else bf16 + hi/lo weights.
usage: one_conv.py fprop|dgrad|wgrad Cin Cout H B [ups] [pn]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200.networks import ops
op, ci, co, H, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
ups, pn = "ups" in sys.argv[6:], "pn" in sys.argv[6:]
hin = H // 2 if ups else H
dt = th.float32 if ops.is_precise(H) else th.bfloat16
x = th.randn(B, ci, hin, hin, device="cuda").to(dt).contiguous(memory_format=th.channels_last)
dy = th.randn(B, co, H, H, device="cuda").to(dt).contiguous(memory_format=th.channels_last)
w = th.randn(co, ci, 3, 3, device="cuda").requires_grad_(True)
bias = th.randn(co, device="cuda")
for _ in range(6):
    if op == "fprop":
        ops.conv3x3(x, w, bias, lrelu=True, upsample_in=ups, pixelnorm=pn, split_w=True, exact_w=not pn)
    elif op == "dgrad":
        ops.conv3x3(dy, w, None, dgrad=True)
    else:
        ops.conv3x3_wgrad(dy, x, upsample_in=ups)
th.cuda.synchronize()
print("ok")
