"""Forward transform throughput as a function of the number of clips processed per launch group (L2 residency of
the per-clip scratch vs launch overhead).  Development aid, GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200 import audio, _lib
from oracle import cases
N = 2_646_000
total = 64
wav = cases.batch_wavs(total, N, seed=1).cuda()
for g in [1, 2, 4, 8, 16, 64]:
    plan = audio.ForwardPlan(N, g, 1)
    def run():
        for i in range(0, total, g):
            plan.run(wav[i:i + g])
    for _ in range(2): run()
    th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): run()
    b.record(); th.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"group {g:3d}: {ms:7.3f} ms per 64 clips -> {total * plan.T / ms / 1e3:8.1f} M frames/s")
