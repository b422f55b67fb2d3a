"""Development aid: GPU time of one convolution shape with parts of the kernel switched off (MG_CONV_ABLATE bit mask:
1 no halo copies, 2 no MMAs, 4 no epilogue work, 8 no global stores, 16 TMEM loads but no arithmetic / stores).
Times come from the library's event profiler (host launch overhead excluded).
usage: ablate_conv.py fprop|dgrad Cin Cout H B     (MASKS=0,1,... to choose)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200 import _lib
from musicgan_b200.networks import ops
op, ci, co, H, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
x = th.randn(B, ci, H, H, device="cuda").bfloat16().contiguous(memory_format=th.channels_last)
if op == "fprop":
    w = th.randn(co, ci, 3, 3, device="cuda").requires_grad_(True)
    bias = th.randn(co, device="cuda")
    fn = lambda: ops.conv3x3(x, w, bias, lrelu=True)
elif op == "dgrad":       # x plays dY (ci channels); the forward weight is (ci, co, 3, 3); the result has co channels
    w = th.randn(ci, co, 3, 3, device="cuda").requires_grad_(True)
    fn = lambda: ops.conv3x3(x, w, None, dgrad=True)
else:                     # wgrad: MG_WGRAD_ABLATE masks (1 no copies, 2 no MMAs, 4 no operand staging)
    dy = th.randn(B, co, H, H, device="cuda").bfloat16().contiguous(memory_format=th.channels_last)
    fn = lambda: ops.conv3x3_wgrad(dy, x)
out = []
for mask in [int(m) for m in os.environ.get("MASKS", "0,1,2,4,8,16,3,5,6,7").split(",")]:
    os.environ["MG_WGRAD_ABLATE" if op == "wgrad" else "MG_CONV_ABLATE"] = str(mask)
    for _ in range(3):
        fn()
    th.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(20):
        fn()
    th.cuda.synchronize()
    prof = _lib.profile_collect(64)
    _lib.profile_enable(False)
    t = sum(v[0] for k, v in prof.items() if k.startswith("k_conv3x3")) / 20 * 1e3      # (wgrad: without the reduce kernel)
    out.append(f"{mask}:{t:.1f}")
print(op, ci, co, H, B, " ".join(out))
