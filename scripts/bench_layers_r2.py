"""Per-layer device time of the convolution kernels at the stage-7 shapes, as the networks run them in round 2
(fp32 split-operand kernel for output height <= ops.PRECISE_MAX_RES, bf16 kernel with hi + lo weights above).  Each
(layer, op) is captured as a CUDA graph of 20 launches so that host launch overhead is not timed.
    python scripts/bench_layers_r2.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200.networks import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
HBM, TF = 6534.8e9, 1358.5e12
G = [(32, 128), (128, 112), (112, 96), (96, 80), (80, 64), (64, 48), (48, 32), (32, 16)]
D = [(16, 32), (32, 48), (48, 64), (64, 80), (80, 96), (96, 112), (112, 128), (128, 144), (144, 160)]
LAYERS = []     # name, Cin, Cout, H(out), upsample_in, pixelnorm
for i, (ci, co) in enumerate(G):
    LAYERS.append((f"G{i}.c1", ci, ci, 2 * 2 ** i, False, True))
    LAYERS.append((f"G{i}.c2", ci, co, 4 * 2 ** i, True, True))
for j, (ci, co) in enumerate(D):
    LAYERS.append((f"D{j}.c1", ci, co, 512 >> j, False, False))
    LAYERS.append((f"D{j}.c2", co, co, 256 >> j, False, False))


def timeit(fn, reps=20):
    fn(); th.cuda.synchronize()
    ws = {}
    g = th.cuda.CUDAGraph()
    with ops.capture_workspaces(ws), th.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); th.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


print(f"batch {B}   {'layer':8s} {'path':6s} {'op':6s} {'us':>8s} {'GB/s':>8s} {'TF/s':>7s} {'%bound':>7s}")
tot = {}
for name, ci, co, H, ups, pn in LAYERS:
    hin = H // 2 if ups else H
    precise = ops.is_precise(H)
    dt = th.float32 if precise else th.bfloat16
    es = 4 if precise else 2
    x = th.randn(B, ci, hin, hin, device="cuda").to(dt).contiguous(memory_format=th.channels_last)
    dy = th.randn(B, co, H, H, device="cuda").to(dt).contiguous(memory_format=th.channels_last)
    w = th.randn(co, ci, 3, 3, device="cuda").requires_grad_(True)
    bias = th.randn(co, device="cuda")
    flops = 2.0 * B * H * H * 9 * ci * co
    by_in, by_out = B * hin * hin * ci * es, B * H * H * co * es
    cases = [("fprop", lambda: ops.conv3x3(x, w, bias, lrelu=True, upsample_in=ups, pixelnorm=pn, split_w=True, exact_w=not pn), by_in + by_out)]
    if not ups:
        cases.append(("dgrad", lambda: ops.conv3x3(dy, w, None, dgrad=True), by_out + B * H * H * ci * es))
    cases.append(("wgrad", lambda: ops.conv3x3_wgrad(dy, x, upsample_in=ups), (by_in + by_out) // (es // 2)))
    for op, fn, byts in cases:
        t = timeit(fn)
        bound = max(flops / TF, byts / HBM)
        tot[(op, precise)] = tot.get((op, precise), 0.0) + t
        print(f"           {name:8s} {'split' if precise else 'bf16':6s} {op:6s} {t * 1e6:8.1f} {byts / t / 1e9:8.0f} {flops / t / 1e12:7.1f} {100 * bound / t:6.1f}%")
for k, v in sorted(tot.items()):
    print(f"sum {k[0]:6s} {'split' if k[1] else 'bf16'}: {v * 1e6:8.1f} us")
