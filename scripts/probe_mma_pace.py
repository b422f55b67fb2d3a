"""Cycles per M128 x N x K16 bf16 tcgen05.mma issued by one thread (debug library probe mg_debug_mma_pace)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200 import build
l = ctypes.CDLL(build.DEBUG_LIB)
l.mg_debug_mma_pace.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p]
out = th.zeros(1, dtype=th.int64, device="cuda")
n_mma = 256
print("cycles per MMA (chain of %d)  mode 0 smem aligned / 1 smem 160-byte groups / 2 A in TMEM; +4: issued warp-convergently (elect.sync)" % n_mma)
for mode in (8,):
    for N in (32, 64, 128, 256):
        row = []
        for n_acc in (1, 2, 4):
            if n_acc * N > 448:
                row.append("   -  "); continue
            for _ in range(2):
                rc = l.mg_debug_mma_pace(out.data_ptr(), N, n_mma, n_acc, mode, None)
                assert rc == 0, rc
                th.cuda.synchronize()
            row.append(f"{out.item() / n_mma:6.1f}")
        print(f"mode {mode} N {N:3d}:  n_acc 1/2/4 = {' '.join(row)}")

print("mode 10: a commit after every P MMAs (two accumulators in rotation); cycles per MMA")
for N in (32, 64, 128):
    row = []
    for period in (1, 2, 5, 10, 20, 40):
        for _ in range(2):
            assert l.mg_debug_mma_pace(out.data_ptr(), N, 240, period, 10, None) == 0
            th.cuda.synchronize()
        row.append(f"P={period}: {out.item() / 240:6.1f}")
    print(f"N {N:3d}: " + "  ".join(row))

print("co-resident CTAs sharing one tensor pipe: per-CTA cycles per MMA (median), SM-level cycles per MMA")
l.mg_debug_mma_pace_grid.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 3 + [ctypes.c_void_p]
sms = th.cuda.get_device_properties(0).multi_processor_count
for N in (32, 64):
    for per_sm in (1, 2, 3, 4):
        ctas = sms * per_sm
        buf = th.zeros(ctas, dtype=th.int64, device="cuda")
        for _ in range(2):
            assert l.mg_debug_mma_pace_grid(buf.data_ptr(), N, 512, ctas, None) == 0
            th.cuda.synchronize()
        med = buf.double().median().item() / 512
        print(f"N {N:3d}  {per_sm} CTA/SM: per-CTA {med:6.1f}  -> per SM {med / per_sm:6.1f}   (max {buf.max().item() / 512:6.1f})")

print("MMA chain while four warps drain other accumulator columns (tcgen05.ld 32x32b.x16 in a loop):")
l.mg_debug_mma_vs_drain.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 3 + [ctypes.c_void_p]
buf = th.zeros(4, dtype=th.int64, device="cuda")
for N in (32, 64, 128):
    for ld_cols in (0, 64, 128):
        for _ in range(2):
            buf.zero_()
            assert l.mg_debug_mma_vs_drain(buf.data_ptr(), N, 1024, ld_cols, None) == 0
            th.cuda.synchronize()
        t, n = buf[0].item(), buf[1].item()
        rate = n * 4 * 32 * 16 * 4 / t if t else 0.0
        print(f"N {N:3d} drain {ld_cols:3d} cols: {t / 1024:6.1f} cycles per MMA, drained {rate:6.1f} B/clk")
