"""Development aid: time one split-operand convolution with parts of the kernel switched off (MG_SPLIT_ABLATE bits:
1 no halo loads / conversion, 2 no weight copies, 4 no MMAs, 8 no stores).  usage: ablate_split.py op Cin Cout H B"""
import os, sys, subprocess
if len(sys.argv) > 6:      # child
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch as th
    from musicgan_b200.networks import ops
    op, ci, co, H, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    x = th.randn(B, ci, H, H, device="cuda").contiguous(memory_format=th.channels_last)
    dy = th.randn(B, co, H, H, device="cuda").contiguous(memory_format=th.channels_last)
    w = th.randn(co, ci, 3, 3, device="cuda").requires_grad_(True)
    bias = th.randn(co, device="cuda")
    fn = (lambda: ops.conv3x3(x, w, bias, lrelu=True, exact_w=True)) if op == "fprop" else (lambda: ops.conv3x3(dy, w, None, dgrad=True))
    fn(); th.cuda.synchronize()
    ws = {}
    g = th.cuda.CUDAGraph()
    with ops.capture_workspaces(ws), th.cuda.graph(g):
        for _ in range(20):
            fn()
    g.replay(); th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); th.cuda.synchronize()
    print(f"{a.elapsed_time(b) / 20 * 1e3:7.1f} us")
else:
    for abl in (0, 1, 2, 4, 8, 3, 7, 15):
        env = dict(os.environ, MG_SPLIT_ABLATE=str(abl))
        r = subprocess.run([sys.executable, __file__] + sys.argv[1:6] + ["child"], env=env, capture_output=True, text=True)
        print(f"{' '.join(sys.argv[1:6]):28s} ablate {abl:2d}: {r.stdout.strip()} {r.stderr.strip()[-120:] if r.returncode else ''}")
    for pdl in ("0",):
        env = dict(os.environ, MG_PDL=pdl)
        r = subprocess.run([sys.executable, __file__] + sys.argv[1:6] + ["child"], env=env, capture_output=True, text=True)
        print(f"{' '.join(sys.argv[1:6]):28s} MG_PDL={pdl}: {r.stdout.strip()}")
