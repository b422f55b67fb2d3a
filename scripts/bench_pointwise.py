"""Development aid (GPU box): GPU time and achieved HBM bandwidth of the memory-bound helper kernels at the stage-7
shapes (times from the library's event profiler)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200 import _lib
from musicgan_b200.networks import ops

HBM = 6534.8e9


def act(B, C, H):
    return th.randn(B, C, H, H, device="cuda").bfloat16().contiguous(memory_format=th.channels_last)


def run(name, fn, nbytes, n=10):
    for _ in range(3):
        fn()
    th.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(n):
        fn()
    th.cuda.synchronize()
    prof = _lib.profile_collect(64)
    _lib.profile_enable(False)
    t = sum(v[0] for v in prof.values()) / n * 1e-3
    print(f"{name:34s} {t * 1e6:8.1f} us {nbytes / t / 1e9:8.0f} GB/s {100 * nbytes / t / HBM:5.1f}% of HBM   [{', '.join(prof)}]")


for B, C, H in [(16, 32, 512), (8, 32, 512), (16, 32, 256), (16, 48, 256), (16, 48, 128), (16, 64, 64), (16, 160, 4)]:
    gy, y = act(B, C, H), act(B, C, H)
    n = gy.numel() * 2
    run(f"lrelu_bwd        B{B} C{C} {H}^2", lambda: ops.lrelu_bwd(gy, y), 3 * n)
    run(f"lrelu_bwd nobias B{B} C{C} {H}^2", lambda: ops.lrelu_bwd(gy, y, want_bias_grad=False), 3 * n)
    if H >= 8:
        run(f"pool2 avg        B{B} C{C} {H}^2", lambda: ops.pool2(y), n + n // 4)
        small = act(B, C, H // 2)
        run(f"pool2 adjoint -> B{B} C{C} {H}^2", lambda: ops.pool2(small, adjoint=True), n + n // 4)
for B, C, H in [(16, 16, 512), (8, 16, 512), (16, 32, 256)]:
    x = th.randn(B, 2, H, H, device="cuda")
    w = th.randn(C, 2, device="cuda"); b = th.randn(C, device="cuda")
    a = act(B, C, H)
    n = a.numel() * 2
    run(f"rgb_expand       B{B} C{C} {H}^2", lambda: ops.rgb_expand(x, w, b, lrelu=True), n + x.numel() * 4)
    run(f"rgb_project      B{B} C{C} {H}^2", lambda: ops.rgb_project(a, w, w_is_c_by_2=True), n + x.numel() * 4)
    run(f"rgb_wgrad        B{B} C{C} {H}^2", lambda: ops.rgb_wgrad(a, a, x), 2 * n + x.numel() * 4)
