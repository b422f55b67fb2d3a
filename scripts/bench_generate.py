"""Config 5 pieces on one GPU: batched generator inference at full depth on a wide latent + the inverse transform
(development aid; prints ms and frames/s)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200 import audio, networks, _lib
nb_vec = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
th.manual_seed(0)
gen = networks.Generator(32, end_layer=7).eval().cuda()
z = th.randn(n, 32, 2, 2 * nb_vec, device="cuda")
def timeit(fn, reps=5):
    for _ in range(2): fn()
    th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); th.cuda.synchronize()
    return a.elapsed_time(b) / reps, out
with th.no_grad():
    ms_g, img = timeit(lambda: gen(z, 1.0))
    print(f"G forward: {n} clips x {512 * nb_vec} frames: {ms_g:.2f} ms  ({n * 8.11 * nb_vec / ms_g:.1f} TFLOP/s algorithmic)")
    _lib.profile_enable(True)
    ms_i, wav = timeit(lambda: audio.magn_phase_to_wave_batch(img, 1))
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    frames = n * 512 * nb_vec
    print(f"inverse: {ms_i:.2f} ms -> {frames / ms_i / 1e3:.1f} M frames/s = {frames * 5120 / ms_i / 1e6:.0f} GB/s algorithmic; "
          f"kernels (ms/launch): " + ", ".join(f"{k} {v[0] / v[1]:.3f}" for k, v in prof.items()))
    print("wav", tuple(wav.shape), float(wav.abs().max()))
