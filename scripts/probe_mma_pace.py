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
for mode in (0, 4, 8, 9):
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
