"""A few launches of the forward transform on 8 x 60 s clips (target for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200 import audio
from oracle import cases
N = 2_646_000
wav = cases.batch_wavs(8, N, seed=1).cuda()
plan = audio.ForwardPlan(N, 8, 1)
for _ in range(4):
    plan.run(wav)
th.cuda.synchronize()
print("ok")
