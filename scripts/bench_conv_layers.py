"""Per-layer timing of the convolution kernels at the stage-7 shapes (development aid, GPU box only):
time, algorithmic HBM bytes -> GB/s, FLOPs -> TFLOP/s, and the per-layer bound min(peak, AI x HBM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200.networks import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
HBM, TF = 6534.8e9, 1358.5e12
LAYERS = [  # name, Cin, Cout, H(out), upsample_in
    ("D0.c1", 16, 32, 512, False), ("D0.c2", 32, 32, 256, False), ("D1.c1", 32, 48, 256, False), ("D1.c2", 48, 48, 128, False),
    ("D2.c1", 48, 64, 128, False), ("D2.c2", 64, 64, 64, False), ("D3.c1", 64, 80, 64, False), ("D4.c1", 80, 96, 32, False),
    ("D6.c1", 112, 128, 8, False), ("D8.c2", 160, 160, 1, False),
    ("G7.c2", 32, 16, 512, True), ("G7.c1", 32, 32, 256, False), ("G6.c2", 48, 32, 256, True), ("G6.c1", 48, 48, 128, False),
    ("G5.c2", 64, 48, 128, True), ("G4.c2", 80, 64, 64, True), ("G1.c1", 128, 128, 4, False),
]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); th.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3


print(f"{'layer':8s} {'op':6s} {'us':>8s} {'GB/s':>8s} {'TF/s':>7s} {'%bound':>7s}")
for name, ci, co, H, ups in LAYERS:
    hin = H // 2 if ups else H
    x = th.randn(B, ci, hin, hin, device="cuda").bfloat16().contiguous(memory_format=th.channels_last)
    dy = th.randn(B, co, H, H, device="cuda").bfloat16().contiguous(memory_format=th.channels_last)
    w = th.randn(co, ci, 3, 3, device="cuda").requires_grad_(True)
    bias = th.randn(co, device="cuda")
    flops = 2.0 * B * H * H * 9 * ci * co
    by_in, by_out = B * hin * hin * ci * 2, B * H * H * co * 2
    cases = [("fprop", lambda: ops.conv3x3(x, w, bias, lrelu=True, upsample_in=ups), by_in + by_out)]
    if not ups:
        cases.append(("dgrad", lambda: ops.conv3x3(dy, w, None, dgrad=True), by_out + B * H * H * ci * 2))
    cases.append(("wgrad", lambda: ops.conv3x3_wgrad(dy, x, upsample_in=ups), by_in + by_out))
    for op, fn, byts in cases:
        t = timeit(fn)
        bound = max(flops / TF, byts / HBM)
        print(f"{name:8s} {op:6s} {t * 1e6:8.1f} {byts / t / 1e9:8.0f} {flops / t / 1e12:7.1f} {100 * bound / t:6.1f}%")
